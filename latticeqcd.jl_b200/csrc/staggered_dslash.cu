// staggered_dslash.cu -- staggered (Kogut-Susskind) Dslash, fp64, sm_100a.
//
// Replaces LinearAlgebra.mul!(y, D::Staggered_Dirac_operator, x) / mul!(y, D', x) of
// LatticeDiracOperators.jl (upstream Dx! + mass), reached from
// src/system/universe.jl:106-110 ("Staggered") through the same call sites as the Wilson operator:
//
//     y(n) = m x(n) +- sum_mu (1/2) eta_mu(n) [ U_mu(n) x(n+mu) - U_mu^dag(n-mu) x(n-mu) ]       (+: D, -: D^dag)
//     eta_1 = 1, eta_2 = (-1)^x, eta_3 = (-1)^(x+y), eta_4 = (-1)^(x+y+z)   (GLOBAL coordinates)
//
// Same mapping as the Wilson kernel: one thread per site, one warp per AoSoA-32 block, 128-bit coalesced
// loads.  672 B/site compulsory (576 B of links), 582 flop/site: purely HBM-bound.
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "site_map.cuh"
#include "halo_pack.cuh"
#include "link_load.cuh"
#include <cstring>

struct StagArgs {
    cplx *out;
    const cplx *in;
    const cplx *gauge;
    const cplx *links12;   // two-row links (G12 kernels, links12.cu), else unused
    Geom g;
    double mass;
    double sign;      // +1 D, -1 D^dag
    double bc[4];
    DslashFuse fuse;
    Reduce red;
    HaloIn halo;      // MULTI kernel only
    HaloOut hout;     // MULTI == 2 (self-packing) only
};

template <int MU, int FWD, int G12>
__device__ __forceinline__ void shop(cplx (&acc)[3], const cplx *__restrict__ in, const cplx *__restrict__ gauge,
                                     int ns, int ls, double coef) {
    const cplx *sp = in + (size_t)(ns >> 5) * (3 * 32) + (ns & 31);
    cplx u[9];
    load_link<MU, G12, 0>(u, gauge, ls);
    cplx v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) v[c] = cscale(coef, ldg128(sp + c * 32));
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = 0; b < 3; b++) {
            if (FWD) cfma(acc[a], u[a * 3 + b], v[b]);
            else     cfmac(acc[a], u[b * 3 + a], v[b]);
        }
    }
}

// off-rank hop (multi-GPU): forward: chi(n+mu) arrives raw, U_mu(n) applied here; backward: U^dag chi(n-mu) arrives complete
template <int MU, int FWD>
__device__ __forceinline__ void halo_shop(cplx (&acc)[3], const StagArgs &A, int s, int f, double eta) {
    const cplx *src = A.halo.recv[MU][FWD] + (size_t)(f >> 5) * (6 * 32) + (f & 31);
    cplx h[3];
#pragma unroll
    for (int c = 0; c < 3; c++) h[c] = __ldcg(src + c * 32);
    if (FWD) {
        const double cf = 0.5 * eta * (A.halo.plast[MU] ? A.bc[MU] : 1.0);
        const cplx *lk = A.gauge + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cplx g = cmake(0.0, 0.0);
#pragma unroll
            for (int b = 0; b < 3; b++) cfma(g, ldg128(lk + (a * 3 + b) * 32), h[b]);
            acc[a].x = fma(cf, g.x, acc[a].x); acc[a].y = fma(cf, g.y, acc[a].y);
        }
    } else {
        const double cf = -0.5 * eta * (A.halo.pfirst[MU] ? A.bc[MU] : 1.0);
#pragma unroll
        for (int c = 0; c < 3; c++) { acc[c].x = fma(cf, h[c].x, acc[c].x); acc[c].y = fma(cf, h[c].y, acc[c].y); }
    }
}

template <int MU, int MULTI, int G12>
__device__ __forceinline__ void shop_pair(cplx (&acc)[3], const StagArgs &A, int s, int coord, int dim, int stride, double eta,
                                          int x, int y, int z, int t) {
    const cplx *links = G12 ? A.links12 : A.gauge;
    {
        bool w = (coord == dim - 1);
        int ns = w ? s - (dim - 1) * stride : s + stride;
        double coef = 0.5 * eta * (w ? A.bc[MU] : 1.0);
        if (!(w && A.g.part[MU])) shop<MU, 1, G12>(acc, A.in, links, ns, s, coef);
        else if (MULTI) halo_shop<MU, 1>(acc, A, s, face_index<MU>(A.g, x, y, z, t), eta);
    }
    {
        bool w = (coord == 0);
        int ns = w ? s + (dim - 1) * stride : s - stride;
        double coef = -0.5 * eta * (w ? A.bc[MU] : 1.0);
        if (!(w && A.g.part[MU])) shop<MU, 0, G12>(acc, A.in, links, ns, ns, coef);
        else if (MULTI) halo_shop<MU, 0>(acc, A, s, face_index<MU>(A.g, x, y, z, t), eta);
    }
}

template <int MULTI, int G12>
__global__ void __launch_bounds__(256) staggered_dslash_kernel(const StagArgs A) {
    if (A.fuse.use_state && A.red.st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int bid = blockIdx.x, npack = 0;
    if (MULTI == 2) {
        npack = A.hout.cta0[4];
        if (bid < npack) { halo_pack_cta(A.g, LQCD_STAGGERED, 0, A.in, A.gauge, A.hout, bid); return; }
        bid -= npack;
    }
    int cta = bid + A.fuse.cta_off;
    unsigned long long ts0 = 0ull;
    if (MULTI) {
        cta = A.halo.cta_order[bid];
        if (A.halo.timing && threadIdx.x == 0) ts0 = global_ns();
        if (bid >= A.halo.n_interior) wait_halo_flags(A.g, A.halo);
    }
    const int blk = block_of_warp(A.g, cta, warp);
    const bool active = blk < A.g.nblk;
    double red[3] = {0.0, 0.0, 0.0};
    if (active) {
        const int s = blk * 32 + lane;
        int x, y, z, t;
        site_coords(A.g, s, x, y, z, t);
        const int gx = x + A.g.o[0], gy = y + A.g.o[1], gz = z + A.g.o[2];
        cplx acc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) acc[k] = cmake(0.0, 0.0);
        shop_pair<0, MULTI, G12>(acc, A, s, x, A.g.X, 1, 1.0, x, y, z, t);
        shop_pair<1, MULTI, G12>(acc, A, s, y, A.g.Y, A.g.X, (gx & 1) ? -1.0 : 1.0, x, y, z, t);
        shop_pair<2, MULTI, G12>(acc, A, s, z, A.g.Z, A.g.X * A.g.Y, ((gx + gy) & 1) ? -1.0 : 1.0, x, y, z, t);
        shop_pair<3, MULTI, G12>(acc, A, s, t, A.g.T, A.g.X * A.g.Y * A.g.Z, ((gx + gy + gz) & 1) ? -1.0 : 1.0, x, y, z, t);
        const size_t base = (size_t)blk * (3 * 32) + lane;
        cplx *dst = A.fuse.axpy_r ? A.fuse.axpy_r : A.out;
        const double malpha = A.fuse.axpy_r ? -A.red.st->alpha : 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            cplx xi = ldg128(A.in + base + k * 32);
            cplx yk = cmake(fma(A.sign, acc[k].x, A.mass * xi.x), fma(A.sign, acc[k].y, A.mass * xi.y));
            if (A.fuse.shift_src) {
                cplx sv = ldg128(A.fuse.shift_src + base + k * 32);
                yk.x = fma(A.fuse.shift, sv.x, yk.x); yk.y = fma(A.fuse.shift, sv.y, yk.y);
            }
            if (A.fuse.axpy_r) {
                cplx rv = A.fuse.axpy_r[base + k * 32];
                yk = cmake(fma(malpha, yk.x, rv.x), fma(malpha, yk.y, rv.y));
            }
            if (A.fuse.dot_with) {
                cplx w = ldg128(A.fuse.dot_with + base + k * 32);
                red[0] = fma(w.x, yk.x, red[0]); red[0] = fma(w.y, yk.y, red[0]);
                red[1] = fma(w.x, yk.y, red[1]); red[1] = fma(-w.y, yk.x, red[1]);
            }
            red[2] = fma(yk.x, yk.x, red[2]); red[2] = fma(yk.y, yk.y, red[2]);
            dst[base + k * 32] = yk;
        }
    }
    if (MULTI && A.halo.timing && threadIdx.x == 0) stamp_span(A.halo.timing, bid >= A.halo.n_interior ? 2 : 1, ts0, global_ns());
    if (A.fuse.dot_with || A.fuse.want_norm)
        grid_reduce_finish<3>(red, A.red, A.fuse.finish, 0, 0, 1, (unsigned)bid, gridDim.x - (unsigned)npack);
}

int launch_staggered_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger,
                            const DslashFuse *fuse, cudaStream_t s, const HaloIn *halo, const HaloOut *hout) {
    if (x == y) return lqcd_fail(ctx, LQCD_ERR_ARG, "dslash: in-place application is not allowed");
    StagArgs A;
    A.out = y; A.in = x; A.gauge = ctx->gauge; A.g = ctx->g; A.mass = op->mass; A.sign = dagger ? -1.0 : 1.0;
    for (int i = 0; i < 4; i++) A.bc[i] = op->bc[i];
    if (fuse) A.fuse = *fuse; else A.fuse = DslashFuse();
    A.red = ctx->red;
    if (halo) A.halo = *halo; else memset(&A.halo, 0, sizeof A.halo);
    if (hout) A.hout = *hout; else memset(&A.hout, 0, sizeof A.hout);
    const int bs = 32 * ctx->g.wpc;
    const bool sub = A.fuse.cta_count > 0;               // slab launch (host_pipeline.cu): single rank, plain epilogue only
    if (sub && (halo || A.fuse.dot_with || A.fuse.want_norm || A.fuse.axpy_r)) return lqcd_fail(ctx, LQCD_ERR_ARG, "sub-range Dslash launch: no halo / reductions");
    const int grid = sub ? A.fuse.cta_count : (ctx->g.nblk + ctx->g.wpc - 1) / ctx->g.wpc + (hout ? hout->cta0[4] : 0);
    int g12 = 0;                                         // two-row links when the links are SU(3) (links12.cu): 480 instead of 672 B/site
    LQCD_TRY(ensure_links12(ctx, &g12));
    A.links12 = g12 ? ctx->links12 : nullptr;
    if (g12) {
        if (halo && hout) staggered_dslash_kernel<2, 1><<<grid, bs, 0, s>>>(A);
        else if (halo) staggered_dslash_kernel<1, 1><<<grid, bs, 0, s>>>(A);
        else      staggered_dslash_kernel<0, 1><<<grid, bs, 0, s>>>(A);
    } else {
        if (halo && hout) staggered_dslash_kernel<2, 0><<<grid, bs, 0, s>>>(A);
        else if (halo) staggered_dslash_kernel<1, 0><<<grid, bs, 0, s>>>(A);
        else      staggered_dslash_kernel<0, 0><<<grid, bs, 0, s>>>(A);
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}
