// wilson_dslash.cu -- Wilson Dslash, fp64, sm_100a.
//
// Replaces LinearAlgebra.mul!(y, D::Wilson_Dirac_operator, x) / mul!(y, D', x) of LatticeDiracOperators.jl
// (upstream Wx!/Wdagx!), the inner kernel of every solve the reference triggers
// (src/md/AbstractMD.jl:129, src/updates/standardHMC.jl:69-71, measure_Pion_correlator.jl:379,399):
//
//     y(n) = x(n) - kappa * sum_mu [ (1 - g_mu) U_mu(n) x(n+mu) + (1 + g_mu) U_mu^dag(n-mu) x(n-mu) ]      (r = 1)
//
// One thread per site, one warp per 32-site block of the AoSoA-32 layout (lqcd_internal.cuh): every
// component load is a coalesced 128-bit-per-lane, 512-byte warp request.  The (1 -+ g_mu) spin projectors
// are applied BEFORE the SU(3) multiply (rank-2: only two colour-vectors per direction are multiplied,
// the other two spin rows are reconstructed by a phase), everything stays in registers, and the xpay,
// the optional sigma-shift and the optional <w,y>, |y|^2 reductions are fused into the epilogue.
// HBM-bound: 960 B/site compulsory, 1368 flop/site -- tensor cores do not apply.
#include "wilson_kernel.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

int launch_wilson_tmarch(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, cudaStream_t s, bool halo, bool self_pack);   // wilson_tmarch.cu

int launch_wilson_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger,
                         const DslashFuse *fuse, cudaStream_t s, const HaloIn *halo, const HaloOut *hout) {
    if (op->r != 1.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "Wilson kernel implements r = 1 only (got r = %g)", op->r);
    if (x == y) return lqcd_fail(ctx, LQCD_ERR_ARG, "dslash: in-place application is not allowed");
    WilsonArgs A;
    A.out = y; A.in = x; A.gauge = ctx->gauge; A.g = ctx->g; A.kappa = op->kappa;
    for (int i = 0; i < 4; i++) A.bc[i] = op->bc[i];
    if (fuse) A.fuse = *fuse; else { A.fuse = DslashFuse(); }
    A.red = ctx->red;
    if (halo) A.halo = *halo; else memset(&A.halo, 0, sizeof A.halo);
    if (hout) A.hout = *hout; else memset(&A.hout, 0, sizeof A.hout);
    A.clover = nullptr; A.links12 = nullptr;
    if (op->csw != 0.0) { LQCD_TRY(ensure_clover(ctx, op)); A.clover = ctx->clover; }
    const int bs = 32 * ctx->g.wpc;
    const bool sub = A.fuse.cta_count > 0;               // slab launch (host_pipeline.cu): single rank, plain epilogue only
    if (sub && (halo || A.fuse.dot_with || A.fuse.want_norm || A.fuse.axpy_r)) return lqcd_fail(ctx, LQCD_ERR_ARG, "sub-range Dslash launch: no halo / reductions");
    const int grid = sub ? A.fuse.cta_count : (ctx->g.nblk + ctx->g.wpc - 1) / ctx->g.wpc + (hout ? hout->cta0[4] : 0);
#define WLK(MU_, LH_, G_)                                                                       \
    do {                                                                                        \
        if (dagger) wilson_dslash_kernel<1, MU_, LH_, 0, G_><<<grid, bs, 0, s>>>(A);            \
        else        wilson_dslash_kernel<0, MU_, LH_, 0, G_><<<grid, bs, 0, s>>>(A);            \
    } while (0)
#define WLG(MU_, LH_) do { if (g12) WLK(MU_, LH_, 1); else WLK(MU_, LH_, 0); } while (0)
#define WL()                                                                                    \
    do {                                                                                        \
        const int mu_ = (halo && hout) ? 2 : (halo ? 1 : 0);                                    \
        if (lh) { if (mu_ == 2) WLG(2, 1); else if (mu_ == 1) WLG(1, 1); else WLG(0, 1); }      \
        else    { if (mu_ == 2) WLG(2, 0); else if (mu_ == 1) WLG(1, 0); else WLG(0, 0); }      \
    } while (0)
    static int lh_env = -2;
    if (lh_env == -2) { const char *e = getenv("LQCD_LINK_HINT"); lh_env = e ? (atoi(e) != 0) : -1; }
    const int lh = lh_env >= 0 ? lh_env : (ctx->g.V <= (1 << 18));
    if (bs > 128) return lqcd_fail(ctx, LQCD_ERR_ARG, "LQCD_WPC > 4 is not supported by the Wilson kernel");
    // kernel family: default = one thread per site, register-resident hops (wilson_kernel.cuh); LQCD_WILSON_KERNEL=4 = t-marching
    // kernel with TMA-staged spinor window and link planes (wilson_tmarch.cu, experimental: 300 vs 173 us at 32^4), which falls
    // through to the default when the geometry does not qualify.  (Rounds 1 / 2 also measured a two-lanes-per-site kernel,
    // 222-355 us, a first t-marching kernel with LDG links, 335-490 us, and a persistent tile-queue variant: all removed.)
    static int family = -1;
    if (family < 0) { const char *e = getenv("LQCD_WILSON_KERNEL"); family = e ? atoi(e) : 0; }
    if (A.clover) {                    // Wilson-clover: CLOVER = 1 instantiations live in wilson_clover.cu
        LQCD_TRY(launch_wilson_clover(ctx, A, dagger, (halo && hout) ? 2 : (halo ? 1 : 0), lh, grid, bs, s));
        ctx->launches++;
        return LQCD_OK;
    }
    if ((family == 4 || family == 5) && !sub) {                 // t-marching TMA kernel; LQCD_ERR_STATE = geometry does not qualify -> family 1
        const int rc = launch_wilson_tmarch(ctx, A, dagger, s, halo != nullptr, hout != nullptr);
        if (rc != LQCD_ERR_STATE) return rc;
    }
    // two-row links (links12.cu) when the links are SU(3): 768 instead of 960 B/site from HBM, 1.8 instead of 2.2 KB/site over the
    // L2 -> SM fabric (the binding limit of this kernel: 2.29 GB per 32^4 application at 11 TB/s, profiles/r2a_wilson_k1_ncu_full.csv)
    int g12 = 0;
    LQCD_TRY(ensure_links12(ctx, &g12));
    A.links12 = g12 ? ctx->links12 : nullptr;
    WL();
#undef WL
#undef WLG
#undef WLK
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}
