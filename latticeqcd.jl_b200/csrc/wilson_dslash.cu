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
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "site_map.cuh"
#include "wilson_spin.cuh"
#include "halo_pack.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

int launch_wilson_dslash2(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, cudaStream_t s);   // wilson_dslash2.cu
int launch_wilson_dslash3(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, cudaStream_t s);   // wilson_dslash3.cu (experimental)

// Cache policy of the link loads.  Links have at most one reuse (as the backward link of the +mu neighbour), spinors up
// to nine.  Measured on B200 (tools/quick_bench.py): marking link lines evict-first in L1 helps when the local lattice is
// L2 resident (32.32.16.8: 34.4 -> 31.3 us) and hurts at 32^4 (203 -> 228 us: the backward-link L1 hits are lost and L2
// is already the bottleneck); L1::no_allocate / spinor evict_last variants were slower in both regimes.  So LH = 1 is
// selected only for local volumes <= 2^18 sites (the strong-scaling regime).
template <int LH>
__device__ __forceinline__ cplx ldlink(const cplx *p) {
    if (LH == 1) {
        cplx v;
        asm("ld.global.nc.L1::evict_first.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
        return v;
    }
    return __ldg(p);
}
__device__ __forceinline__ cplx ldspinor(const cplx *p) { return __ldg(p); }

// one of the eight hops.  FWD=1: U_mu(n) x(n+mu) with link at `ls` = n;  FWD=0: U_mu^dag(n-mu) x(n-mu), ls = n-mu.
template <int MU, int FWD, int DAG, int LH>
__device__ __forceinline__ void hop(cplx (&acc)[12], const cplx *__restrict__ in, const cplx *__restrict__ gauge,
                                    int ns, int ls, bool wrapped, double phase) {
    constexpr int S = (FWD ^ DAG) ? -1 : +1;     // D: forward (1-g), backward (1+g); D^dag swaps
    const cplx *sp = in + (size_t)(ns >> 5) * (12 * 32) + (ns & 31);
    cplx h0[3], h1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        cplx p0 = ldspinor(sp + (0 + c) * 32), p1 = ldspinor(sp + (3 + c) * 32);
        cplx p2 = ldspinor(sp + (6 + c) * 32), p3 = ldspinor(sp + (9 + c) * 32);
        project<MU, S>(h0[c], h1[c], p0, p1, p2, p3);
    }
    if (wrapped) {
#pragma unroll
        for (int c = 0; c < 3; c++) { h0[c] = cscale(phase, h0[c]); h1[c] = cscale(phase, h1[c]); }
    }
    const cplx *lk = gauge + ((size_t)(ls >> 5) * 4 + MU) * (9 * 32) + (ls & 31);
#pragma unroll
    for (int a = 0; a < 3; a++) {
        cplx g0 = cmake(0.0, 0.0), g1 = cmake(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            if (FWD) {
                cplx u = ldlink<LH>(lk + (a * 3 + b) * 32);
                cfma(g0, u, h0[b]); cfma(g1, u, h1[b]);
            } else {
                cplx u = ldlink<LH>(lk + (b * 3 + a) * 32);
                cfmac(g0, u, h0[b]); cfmac(g1, u, h1[b]);
            }
        }
        reconstruct<MU, S>(acc, a, g0, g1);
    }
}

// off-rank hops (multi-GPU): the neighbour's pack kernel already delivered the spin-projected half spinor
// (forward hop: P psi(n+mu), U_mu(n) is applied here; backward hop: U^dag P psi(n-mu), complete) into our halo slot.
template <int MU, int FWD, int DAG>
__device__ __forceinline__ void halo_hop(cplx (&acc)[12], const WilsonArgs &A, int s, int f) {
    constexpr int S = (FWD ^ DAG) ? -1 : +1;
    const cplx *src = A.halo.recv[MU][FWD] + (size_t)(f >> 5) * (6 * 32) + (f & 31);
    cplx h0[3], h1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { h0[c] = __ldcg(src + c * 32); h1[c] = __ldcg(src + (3 + c) * 32); }
    const double phase = FWD ? (A.halo.plast[MU] ? A.bc[MU] : 1.0) : (A.halo.pfirst[MU] ? A.bc[MU] : 1.0);
    if (phase != 1.0) {
#pragma unroll
        for (int c = 0; c < 3; c++) { h0[c] = cscale(phase, h0[c]); h1[c] = cscale(phase, h1[c]); }
    }
    if (FWD) {
        const cplx *lk = A.gauge + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cplx g0 = cmake(0.0, 0.0), g1 = cmake(0.0, 0.0);
#pragma unroll
            for (int b = 0; b < 3; b++) {
                cplx u = ldlink<0>(lk + (a * 3 + b) * 32);
                cfma(g0, u, h0[b]); cfma(g1, u, h1[b]);
            }
            reconstruct<MU, S>(acc, a, g0, g1);
        }
    } else {
#pragma unroll
        for (int a = 0; a < 3; a++) reconstruct<MU, S>(acc, a, h0[a], h1[a]);
    }
}

template <int MU, int DAG, int MULTI, int LH>
__device__ __forceinline__ void hop_pair(cplx (&acc)[12], const WilsonArgs &A, int s, int coord, int dim, int stride,
                                         int x, int y, int z, int t) {
    // forward
    {
        bool w = (coord == dim - 1);
        int ns = w ? s - (dim - 1) * stride : s + stride;
        if (!(w && A.g.part[MU])) hop<MU, 1, DAG, LH>(acc, A.in, A.gauge, ns, s, w, A.bc[MU]);
        else if (MULTI) halo_hop<MU, 1, DAG>(acc, A, s, face_index<MU>(A.g, x, y, z, t));
    }
    // backward
    {
        bool w = (coord == 0);
        int ns = w ? s + (dim - 1) * stride : s - stride;
        if (!(w && A.g.part[MU])) hop<MU, 0, DAG, LH>(acc, A.in, A.gauge, ns, ns, w, A.bc[MU]);
        else if (MULTI) halo_hop<MU, 0, DAG>(acc, A, s, face_index<MU>(A.g, x, y, z, t));
    }
}

template <int DAG, int MAXT, int MINB, int MULTI, int LH>
__global__ void __launch_bounds__(MAXT, MINB) wilson_dslash_kernel(const WilsonArgs A) {
    if (A.fuse.use_state && A.red.st->done) return;     // grid-uniform: set only by an earlier kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int bid = blockIdx.x, npack = 0;
    if (MULTI == 2) {      // self-packing: the first npack CTAs ship this application's halo to the neighbours
        npack = A.hout.cta0[4];
        if (bid < npack) { halo_pack_cta(A.g, LQCD_WILSON, DAG, A.in, A.gauge, A.hout, bid); return; }
        bid -= npack;
    }
    int cta = bid;
    if (MULTI) {
        cta = A.halo.cta_order[bid];
        if (bid >= A.halo.n_interior) wait_halo_flags(A.g, A.halo);
    }
    const int blk = block_of_warp(A.g, cta, warp);
    const bool active = blk < A.g.nblk;
    double red[3] = {0.0, 0.0, 0.0};
    if (active) {
        const int s = blk * 32 + lane;
        int x, y, z, t;
        site_coords(A.g, s, x, y, z, t);
        // multi-GPU interior pass: face sites are finished (and reduced) by the exterior kernel
        const bool skip_red = A.fuse.interior_only &&
            ((A.g.part[0] && (x == 0 || x == A.g.X - 1)) || (A.g.part[1] && (y == 0 || y == A.g.Y - 1)) ||
             (A.g.part[2] && (z == 0 || z == A.g.Z - 1)) || (A.g.part[3] && (t == 0 || t == A.g.T - 1)));
        cplx acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = cmake(0.0, 0.0);
        {   // the fused epilogue reads its extra operands ~10 us from now, after the last hop: start them towards L2
            const size_t pb = (size_t)blk * (12 * 32) + lane;
            if (A.fuse.axpy_r || A.fuse.dot_with || A.fuse.shift_src) {
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    if (A.fuse.axpy_r) prefetch_l2(A.fuse.axpy_r + pb + k * 32);
                    if (A.fuse.dot_with) prefetch_l2(A.fuse.dot_with + pb + k * 32);
                    if (A.fuse.shift_src) prefetch_l2(A.fuse.shift_src + pb + k * 32);
                }
            }
        }
        hop_pair<0, DAG, MULTI, LH>(acc, A, s, x, A.g.X, 1, x, y, z, t);
        hop_pair<1, DAG, MULTI, LH>(acc, A, s, y, A.g.Y, A.g.X, x, y, z, t);
        hop_pair<2, DAG, MULTI, LH>(acc, A, s, z, A.g.Z, A.g.X * A.g.Y, x, y, z, t);
        hop_pair<3, DAG, MULTI, LH>(acc, A, s, t, A.g.T, A.g.X * A.g.Y * A.g.Z, x, y, z, t);
        const size_t base = (size_t)blk * (12 * 32) + lane;
        const double mk = -A.kappa;
        cplx *dst = A.fuse.axpy_r ? A.fuse.axpy_r : A.out;
        const double malpha = A.fuse.axpy_r ? -A.red.st->alpha : 0.0;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            cplx xi = ldg128(A.in + base + k * 32);
            cplx yk = cmake(fma(mk, acc[k].x, xi.x), fma(mk, acc[k].y, xi.y));
            if (A.fuse.shift_src) {
                cplx sv = ldg128(A.fuse.shift_src + base + k * 32);
                yk.x = fma(A.fuse.shift, sv.x, yk.x); yk.y = fma(A.fuse.shift, sv.y, yk.y);
            }
            if (A.fuse.axpy_r) {           // fused CG residual update: r <- r - alpha * (D^dag t); q is never stored
                cplx rv = A.fuse.axpy_r[base + k * 32];
                yk = cmake(fma(malpha, yk.x, rv.x), fma(malpha, yk.y, rv.y));
            }
            if (A.fuse.dot_with && !skip_red) {
                cplx w = ldg128(A.fuse.dot_with + base + k * 32);
                red[0] = fma(w.x, yk.x, red[0]); red[0] = fma(w.y, yk.y, red[0]);
                red[1] = fma(w.x, yk.y, red[1]); red[1] = fma(-w.y, yk.x, red[1]);
            }
            if (!skip_red) { red[2] = fma(yk.x, yk.x, red[2]); red[2] = fma(yk.y, yk.y, red[2]); }
            dst[base + k * 32] = yk;
        }
    }
    if (A.fuse.dot_with || A.fuse.want_norm)
        grid_reduce_finish<3>(red, A.red, A.fuse.finish, 0, 0, !A.fuse.interior_only, (unsigned)bid, gridDim.x - (unsigned)npack);
}

int launch_wilson_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger,
                         const DslashFuse *fuse, cudaStream_t s, const HaloIn *halo, const HaloOut *hout) {
    if (op->r != 1.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "Wilson kernel implements r = 1 only (got r = %g)", op->r);
    if (op->csw != 0.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "clover term not built (csw = %g)", op->csw);
    if (x == y) return lqcd_fail(ctx, LQCD_ERR_ARG, "dslash: in-place application is not allowed");
    WilsonArgs A;
    A.out = y; A.in = x; A.gauge = ctx->gauge; A.g = ctx->g; A.kappa = op->kappa;
    for (int i = 0; i < 4; i++) A.bc[i] = op->bc[i];
    if (fuse) A.fuse = *fuse; else { A.fuse = DslashFuse(); }
    A.red = ctx->red;
    if (halo) A.halo = *halo; else memset(&A.halo, 0, sizeof A.halo);
    if (hout) A.hout = *hout; else memset(&A.hout, 0, sizeof A.hout);
    const int bs = 32 * ctx->g.wpc;
    const int grid = (ctx->g.nblk + ctx->g.wpc - 1) / ctx->g.wpc + (hout ? hout->cta0[4] : 0);
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
        if (dagger) wilson_dslash_kernel<1, MT, MB, MU_, LH_><<<grid, bs, 0, s>>>(A);           \
        else        wilson_dslash_kernel<0, MT, MB, MU_, LH_><<<grid, bs, 0, s>>>(A);           \
    } while (0)
#define WL(MT, MB)                                                                              \
    do {                                                                                        \
        const int mu_ = (halo && hout) ? 2 : (halo ? 1 : 0);                                    \
        if (lh) { if (mu_ == 2) WLK(MT, MB, 2, 1); else if (mu_ == 1) WLK(MT, MB, 1, 1); else WLK(MT, MB, 0, 1); } \
        else    { if (mu_ == 2) WLK(MT, MB, 2, 0); else if (mu_ == 1) WLK(MT, MB, 1, 0); else WLK(MT, MB, 0, 0); } \
    } while (0)
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
    if (family == 3 && !halo) {        // experimental t-marching kernel; falls through when the geometry does not qualify
        const int rc = launch_wilson_dslash3(ctx, A, dagger, s);
        if (rc != LQCD_ERR_STATE) return rc;
    }
    if (family == 2 && !halo && bs <= 128 && !A.fuse.axpy_r) return launch_wilson_dslash2(ctx, A, dagger, s);
    // measured on B200, 32^4: 206 regs (8 warps/SM) 236 us; 168 regs (12 warps/SM) 200 us; 128 regs (16 warps/SM,
    // 136 B spills) 204 us -- the kernel is latency bound (ncu: 57% long-scoreboard stalls), so 168 is the default.
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
