// wilson_kernel.cuh -- the Wilson Dslash kernel template (hops, fused epilogue), shared by wilson_dslash.cu (CLOVER = 0
// instantiations, the verified default path) and wilson_clover.cu (CLOVER = 1: y = A x - kappa*hop with the site-local
// clover term A applied in the epilogue).  See wilson_dslash.cu for the design notes.
#pragma once
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "site_map.cuh"
#include "wilson_spin.cuh"
#include "halo_pack.cuh"
#include "link_load.cuh"

__device__ __forceinline__ cplx ldspinor(const cplx *p) { return __ldg(p); }

// one of the eight hops.  FWD=1: U_mu(n) x(n+mu) with link at `ls` = n;  FWD=0: U_mu^dag(n-mu) x(n-mu), ls = n-mu.
template <int MU, int FWD, int DAG, int LH, int G12>
__device__ __forceinline__ void hop(cplx (&acc)[12], const cplx *__restrict__ in, const cplx *__restrict__ gauge,
                                    int ns, int ls, bool wrapped, double phase) {
    constexpr int S = (FWD ^ DAG) ? -1 : +1;     // D: forward (1-g), backward (1+g); D^dag swaps
    const cplx *sp = in + (size_t)(ns >> 5) * (12 * 32) + (ns & 31);
    cplx u[9];
    load_link<MU, G12, LH>(u, gauge, ls);
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
#pragma unroll
    for (int a = 0; a < 3; a++) {
        cplx g0 = cmake(0.0, 0.0), g1 = cmake(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            if (FWD) { cfma(g0, u[a * 3 + b], h0[b]); cfma(g1, u[a * 3 + b], h1[b]); }
            else     { cfmac(g0, u[b * 3 + a], h0[b]); cfmac(g1, u[b * 3 + a], h1[b]); }
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

template <int MU, int DAG, int MULTI, int LH, int G12>
__device__ __forceinline__ void hop_pair(cplx (&acc)[12], const WilsonArgs &A, int s, int coord, int dim, int stride,
                                         int x, int y, int z, int t) {
    const cplx *links = G12 ? A.links12 : A.gauge;
    // forward
    {
        bool w = (coord == dim - 1);
        int ns = w ? s - (dim - 1) * stride : s + stride;
        if (!(w && A.g.part[MU])) hop<MU, 1, DAG, LH, G12>(acc, A.in, links, ns, s, w, A.bc[MU]);
        else if (MULTI) halo_hop<MU, 1, DAG>(acc, A, s, face_index<MU>(A.g, x, y, z, t));
    }
    // backward
    {
        bool w = (coord == 0);
        int ns = w ? s + (dim - 1) * stride : s - stride;
        if (!(w && A.g.part[MU])) hop<MU, 0, DAG, LH, G12>(acc, A.in, links, ns, ns, w, A.bc[MU]);
        else if (MULTI) halo_hop<MU, 0, DAG>(acc, A, s, face_index<MU>(A.g, x, y, z, t));
    }
}

// Site-local clover term: ax = A(n) x(n).  A is block diagonal in chirality (spins 01 | 23); each 6x6 Hermitian block is
// stored packed as 18 complex numbers (AoSoA-32, stride 32): e = 0..2 -> (A00,A11),(A22,A33),(A44,A55) real diagonal pairs,
// e = 3 + i(i-1)/2 + j -> A_ij for i > j (row i, column j; index = 3*spin_in_block + colour).  576 B/site, SURVEY.md 8d.
__device__ __forceinline__ void clover_apply(cplx (&ax)[12], const cplx *__restrict__ cl, const cplx *__restrict__ xin) {
#pragma unroll
    for (int b = 0; b < 2; b++) {
        cplx xv[6], yv[6];
#pragma unroll
        for (int i = 0; i < 6; i++) xv[i] = ldg128(xin + (6 * b + i) * 32);
#pragma unroll
        for (int h = 0; h < 3; h++) {
            const cplx d = ldg128(cl + (18 * b + h) * 32);
            yv[2 * h] = cscale(d.x, xv[2 * h]);
            yv[2 * h + 1] = cscale(d.y, xv[2 * h + 1]);
        }
#pragma unroll
        for (int i = 1; i < 6; i++)
#pragma unroll
            for (int j = 0; j < i; j++) {
                const cplx c = ldg128(cl + (18 * b + 3 + i * (i - 1) / 2 + j) * 32);
                cfma(yv[i], c, xv[j]);          // A_ij x_j
                cfmac(yv[j], c, xv[i]);         // A_ji x_i = conj(A_ij) x_i
            }
#pragma unroll
        for (int i = 0; i < 6; i++) ax[6 * b + i] = yv[i];
    }
}

// MULTI: 0 single GPU | 1 fused halo (separate pack kernel on the priority stream) | 2 self-packing (the leading CTAs of this
// kernel ship the halo).  (A persistent-CTA tile-queue variant was measured in rounds 1 / 2: slower at every N, removed.)
// G12: links read from the two-row copy (links12.cu).  __launch_bounds__(128, 4): 128 registers, 16 warps per SM.  Measured on
// B200 at 32^4 (round 2): with FULL links 12 / 14 / 16 warps per SM ran 202 / 206 / 211 us (profiles/r2a_kernel_sweep.txt) -- the
// kernel sat at the L2 -> SM fabric limit and more warps only lowered the L1 hit rate; with two-row links 12 warps take 182 us and
// 16 warps 173 us (26.6 vs 28.8 us on the 8-GPU local volume; profiles/r2c_occupancy_with_links12.txt).
#ifndef LQCD_WILSON_MINB
#define LQCD_WILSON_MINB 4          // 128 registers, 16 warps per SM (experiment builds: LQCD_BUILD_DEFS=-DLQCD_WILSON_MINB=3 -> 168 registers, 12 warps)
#endif
template <int DAG, int MULTI, int LH, int CLOVER, int G12>
__global__ void __launch_bounds__(128, LQCD_WILSON_MINB) wilson_dslash_kernel(const WilsonArgs A) {
    if (A.fuse.use_state && A.red.st->done) return;     // grid-uniform: set only by an earlier kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int bid = blockIdx.x, npack = 0;
    if (MULTI == 2) {      // self-packing: the first npack CTAs ship this application's halo to the neighbours
        npack = A.hout.cta0[4];
        if (bid < npack) { halo_pack_cta(A.g, LQCD_WILSON, DAG, A.in, A.gauge, A.hout, bid); return; }
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
        cplx acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = cmake(0.0, 0.0);
        {   // the fused epilogue reads its extra operands ~10 us from now, after the last hop: start them towards L2
            const size_t pb = (size_t)blk * (12 * 32) + lane;
            if (CLOVER) {       // packed clover blocks of this site: 36 x 16 B, read only in the epilogue
                const cplx *cl = A.clover + (size_t)blk * (36 * 32) + lane;
#pragma unroll
                for (int e = 0; e < 36; e++) prefetch_l2(cl + e * 32);
            }
            if (A.fuse.axpy_r || A.fuse.dot_with || A.fuse.shift_src) {
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    if (A.fuse.axpy_r) prefetch_l2(A.fuse.axpy_r + pb + k * 32);
                    if (A.fuse.dot_with) prefetch_l2(A.fuse.dot_with + pb + k * 32);
                    if (A.fuse.shift_src) prefetch_l2(A.fuse.shift_src + pb + k * 32);
                }
            }
        }
        hop_pair<0, DAG, MULTI, LH, G12>(acc, A, s, x, A.g.X, 1, x, y, z, t);
        hop_pair<1, DAG, MULTI, LH, G12>(acc, A, s, y, A.g.Y, A.g.X, x, y, z, t);
        hop_pair<2, DAG, MULTI, LH, G12>(acc, A, s, z, A.g.Z, A.g.X * A.g.Y, x, y, z, t);
        hop_pair<3, DAG, MULTI, LH, G12>(acc, A, s, t, A.g.T, A.g.X * A.g.Y * A.g.Z, x, y, z, t);
        const size_t base = (size_t)blk * (12 * 32) + lane;
        const double mk = -A.kappa;
        cplx *dst = A.fuse.axpy_r ? A.fuse.axpy_r : A.out;
        const double malpha = A.fuse.axpy_r ? -A.red.st->alpha : 0.0;
        cplx ax[12];                    // CLOVER only (dead otherwise)
        if constexpr (CLOVER != 0) clover_apply(ax, A.clover + (size_t)blk * (36 * 32) + lane, A.in + base);
#pragma unroll
        for (int k = 0; k < 12; k++) {
            cplx xi;
            if constexpr (CLOVER != 0) xi = ax[k]; else xi = ldg128(A.in + base + k * 32);
            cplx yk = cmake(fma(mk, acc[k].x, xi.x), fma(mk, acc[k].y, xi.y));
            if (A.fuse.shift_src) {
                cplx sv = ldg128(A.fuse.shift_src + base + k * 32);
                yk.x = fma(A.fuse.shift, sv.x, yk.x); yk.y = fma(A.fuse.shift, sv.y, yk.y);
            }
            if (A.fuse.axpy_r) {           // fused CG residual update: r <- r - alpha * (D^dag t); q is never stored
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

