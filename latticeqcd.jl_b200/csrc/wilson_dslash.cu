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

int launch_wilson_dslash2(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, cudaStream_t s);   // wilson_dslash2.cu
int launch_wilson_dslash3(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, cudaStream_t s);   // wilson_dslash3.cu (experimental)

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
    A.clover = nullptr;
    if (op->csw != 0.0) { LQCD_TRY(ensure_clover(ctx, op)); A.clover = ctx->clover; }
    const int bs = 32 * ctx->g.wpc;
    const bool sub = A.fuse.cta_count > 0;               // slab launch (host_pipeline.cu): single rank, plain epilogue only
    if (sub && (halo || A.fuse.dot_with || A.fuse.want_norm || A.fuse.axpy_r)) return lqcd_fail(ctx, LQCD_ERR_ARG, "sub-range Dslash launch: no halo / reductions");
    const int grid = sub ? A.fuse.cta_count : (ctx->g.nblk + ctx->g.wpc - 1) / ctx->g.wpc + (hout ? hout->cta0[4] : 0);
    // register budget variants (tuning knob LQCD_LB = "maxthreads,minblocks"; default picked by measurement)
    static int lb = -1;
    if (lb < 0) {
        lb = 0;
        if (const char *e = getenv("LQCD_LB")) {
            int a = 0, b = 0;
            if (sscanf(e, "%d,%d", &a, &b) == 2) lb = a * 100 + b;
        }
    }
#define WLK(MT, MB, MU_, LH_)                                                                   \
    do {                                                                                        \
        if (dagger) wilson_dslash_kernel<1, MT, MB, MU_, LH_, 0><<<grid, bs, 0, s>>>(A);           \
        else        wilson_dslash_kernel<0, MT, MB, MU_, LH_, 0><<<grid, bs, 0, s>>>(A);           \
    } while (0)
#define WL(MT, MB)                                                                              \
    do {                                                                                        \
        const int mu_ = (halo && hout) ? 2 : (halo ? 1 : 0);                                    \
        if (lh) { if (mu_ == 2) WLK(MT, MB, 2, 1); else if (mu_ == 1) WLK(MT, MB, 1, 1); else WLK(MT, MB, 0, 1); } \
        else    { if (mu_ == 2) WLK(MT, MB, 2, 0); else if (mu_ == 1) WLK(MT, MB, 1, 0); else WLK(MT, MB, 0, 0); } \
    } while (0)
    // experiment (LQCD_PERSIST=1, default off): one wave of persistent CTAs drawing tiles from a queue -- self-packing
    // multi-GPU launches and plain single-GPU launches of the default register budget only
    static int persist = -1;
    if (persist < 0) { const char *e = getenv("LQCD_PERSIST"); persist = (e && atoi(e) == 1) ? 1 : 0; }
    static int lh_env = -2;
    if (lh_env == -2) { const char *e = getenv("LQCD_LINK_HINT"); lh_env = e ? (atoi(e) != 0) : -1; }
    const int lh = lh_env >= 0 ? lh_env : (ctx->g.V <= (1 << 18));
    if (bs > 256) return lqcd_fail(ctx, LQCD_ERR_ARG, "LQCD_WPC > 8 is not supported by the Wilson kernel");
    // kernel family: 1 = one lane per site (this file, default), 2 = two lanes per site (wilson_dslash2.cu).
    // Measured on B200 at 32^4: family 2 with 16/24/32 warps per SM runs 222/261/355 us against 192 us here --
    // more resident warps LOWER the L1 hit rate and the kernel then saturates the ~10.8 TB/s L2->SM fabric
    // (2.1 GB of L2 reads per application at 28 % L1 hits), so occupancy is not the lever; L2 traffic is.
    static int family = -1;
    if (family < 0) { const char *e = getenv("LQCD_WILSON_KERNEL"); family = (e && (atoi(e) == 2 || atoi(e) == 3)) ? atoi(e) : 1; }
    if (A.clover) {                    // Wilson-clover: CLOVER = 1 instantiations live in wilson_clover.cu
        LQCD_TRY(launch_wilson_clover(ctx, A, dagger, (halo && hout) ? 2 : (halo ? 1 : 0), lh, grid, bs, s));
        ctx->launches++;
        return LQCD_OK;
    }
    if (family == 3 && !halo && !sub) {        // experimental t-marching kernel; falls through when the geometry does not qualify
        const int rc = launch_wilson_dslash3(ctx, A, dagger, s);
        if (rc != LQCD_ERR_STATE) return rc;
    }
    if (family == 2 && !halo && !sub && bs <= 128 && !A.fuse.axpy_r) return launch_wilson_dslash2(ctx, A, dagger, s);
    // measured on B200, 32^4: 206 regs (8 warps/SM) 236 us; 168 regs (12 warps/SM) 200 us; 128 regs (16 warps/SM,
    // 136 B spills) 204 us -- the kernel is latency bound (ncu: 57% long-scoreboard stalls), so 168 is the default.
    if (persist && !sub && bs == 128 && lb == 0 && ((halo && hout) || !halo)) {
        const int pgrid = grid < 3 * ctx->num_sms ? grid : 3 * ctx->num_sms;      // __launch_bounds__(128, 3): 3 CTAs per SM
        A.fuse.queue = ctx->queue; A.fuse.queue_total = grid;
#define WLP(MU_, LH_)                                                                           \
    do {                                                                                        \
        if (dagger) wilson_dslash_kernel<1, 128, 3, MU_, LH_, 0><<<pgrid, bs, 0, s>>>(A);       \
        else        wilson_dslash_kernel<0, 128, 3, MU_, LH_, 0><<<pgrid, bs, 0, s>>>(A);       \
    } while (0)
        if (halo) { if (lh) WLP(3, 1); else WLP(3, 0); }
        else      { if (lh) WLP(4, 1); else WLP(4, 0); }
#undef WLP
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
        return LQCD_OK;
    }
    if (lb == 12804 && bs <= 128) WL(128, 4);
    else if (lb == 25602) WL(256, 2);
    else if (lb == 25601 || bs > 128) WL(256, 1);
    else WL(128, 3);
#undef WL
#undef WLK
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}
