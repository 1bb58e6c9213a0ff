// wilson_dslash2.cu -- Wilson Dslash, fp64, sm_100a, TWO LANES PER SITE.
//
// Same operator and layout as wilson_dslash.cu (mul!(y, D, x), upstream Wx!) but every
// site is shared by two lanes of a warp: lane (j, sc) owns half-spinor component sc in {0,1} of the
// spin-projected hop, i.e. it multiplies ONE colour 3-vector per direction by the link and accumulates three
// spin rows (row sc, and the reconstructed rows 3-sc / 2+sc) instead of four.  Per-thread state drops from
// 12 to 9 complex accumulators and from 6 to 3 complex half-spinor entries, so ~2x the warps are resident per
// SM and twice as many independent loads are in flight per site -- the single-lane kernel is latency bound
// (profiles/README.md: 57 % long-scoreboard stalls at 12 warps/SM), not bandwidth bound.
//
// A warp covers 16 sites: lanes 0-15 hold sc = 0, lanes 16-31 hold sc = 1 of the same sites, so a spinor
// component load is two contiguous 256-byte segments and the link loads of the two half-warps coincide
// (one L1 wavefront serves both).  The rows 2,3 partial sums are exchanged with one shfl.xor(16) per
// component at the end.
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "site_map.cuh"
#include <cstdio>
#include <cstdlib>

// lane-dependent sign sigma = +1 (sc = 0), -1 (sc = 1)
template <int MU, int FWD, int DAG>
__device__ __forceinline__ void hop2(cplx (&accA)[3], cplx (&accB)[3], cplx (&accC)[3], const cplx *__restrict__ in,
                                     const cplx *__restrict__ gauge, int ns, int ls, bool wrapped, double phase,
                                     int sc, double sigma) {
    constexpr int S = (FWD ^ DAG) ? -1 : +1;     // D: forward (1-g), backward (1+g); D^dag swaps
    // projection h = psi_a + q * (i or 1) * psi_b  (wilson_spin.cuh table, one component per lane)
    const int row_b = (MU < 2) ? 3 - sc : 2 + sc;
    const double q = (MU == 0 || MU == 3) ? -(double)S : -(double)S * sigma;
    const cplx *spa = in + (size_t)(ns >> 5) * (12 * 32) + (ns & 31) + sc * (3 * 32);
    const cplx *spb = in + (size_t)(ns >> 5) * (12 * 32) + (ns & 31) + row_b * (3 * 32);
    cplx h[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const cplx pa = ldg128(spa + c * 32), pb = ldg128(spb + c * 32);
        if (MU == 0 || MU == 2) h[c] = cmake(fma(-q, pb.y, pa.x), fma(q, pb.x, pa.y));     // pa + q*i*pb
        else                    h[c] = cmake(fma(q, pb.x, pa.x), fma(q, pb.y, pa.y));      // pa + q*pb
    }
    if (wrapped) {
#pragma unroll
        for (int c = 0; c < 3; c++) h[c] = cscale(phase, h[c]);
    }
    const cplx *lk = gauge + ((size_t)(ls >> 5) * 4 + MU) * (9 * 32) + (ls & 31);
    // reconstruction coefficient of the second row this lane feeds:
    //   MU=0: row 3-sc += S i g ; MU=1: row 3-sc += -sigma S g ; MU=2: row 2+sc += sigma S i g ; MU=3: row 2+sc += -S g
    const double rc = (MU == 0) ? (double)S : (MU == 1) ? -(double)S * sigma : (MU == 2) ? (double)S * sigma : -(double)S;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        cplx g = cmake(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            if (FWD) cfma(g, ldg128(lk + (a * 3 + b) * 32), h[b]);
            else     cfmac(g, ldg128(lk + (b * 3 + a) * 32), h[b]);
        }
        accA[a] = cadd(accA[a], g);
        if (MU == 0)      { accB[a].x = fma(-rc, g.y, accB[a].x); accB[a].y = fma(rc, g.x, accB[a].y); }   // += rc*i*g
        else if (MU == 1) { accB[a].x = fma(rc, g.x, accB[a].x);  accB[a].y = fma(rc, g.y, accB[a].y); }   // += rc*g
        else if (MU == 2) { accC[a].x = fma(-rc, g.y, accC[a].x); accC[a].y = fma(rc, g.x, accC[a].y); }
        else              { accC[a].x = fma(rc, g.x, accC[a].x);  accC[a].y = fma(rc, g.y, accC[a].y); }
    }
}

template <int MU, int DAG>
__device__ __forceinline__ void hop2_pair(cplx (&accA)[3], cplx (&accB)[3], cplx (&accC)[3], const WilsonArgs &A, int s,
                                          int coord, int dim, int stride, int sc, double sigma) {
    {
        const bool w = (coord == dim - 1);
        const int ns = w ? s - (dim - 1) * stride : s + stride;
        if (!(w && A.g.part[MU])) hop2<MU, 1, DAG>(accA, accB, accC, A.in, A.gauge, ns, s, w, A.bc[MU], sc, sigma);
    }
    {
        const bool w = (coord == 0);
        const int ns = w ? s + (dim - 1) * stride : s - stride;
        if (!(w && A.g.part[MU])) hop2<MU, 0, DAG>(accA, accB, accC, A.in, A.gauge, ns, ns, w, A.bc[MU], sc, sigma);
    }
}

template <int DAG, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) wilson_dslash2_kernel(const WilsonArgs A) {
    if (A.fuse.use_state && A.red.st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = block_of_warp(A.g, blockIdx.x, warp >> 1);
    const bool active = blk < A.g.nblk;
    const int sc = lane >> 4;
    const int j = (lane & 15) + ((warp & 1) << 4);
    const double sigma = sc ? -1.0 : 1.0;
    double red[3] = {0.0, 0.0, 0.0};
    if (active) {
        const int s = blk * 32 + j;
        int x, y, z, t;
        site_coords(A.g, s, x, y, z, t);
        cplx accA[3], accB[3], accC[3];
#pragma unroll
        for (int k = 0; k < 3; k++) accA[k] = accB[k] = accC[k] = cmake(0.0, 0.0);
        hop2_pair<0, DAG>(accA, accB, accC, A, s, x, A.g.X, 1, sc, sigma);
        hop2_pair<1, DAG>(accA, accB, accC, A, s, y, A.g.Y, A.g.X, sc, sigma);
        hop2_pair<2, DAG>(accA, accB, accC, A, s, z, A.g.Z, A.g.X * A.g.Y, sc, sigma);
        hop2_pair<3, DAG>(accA, accB, accC, A, s, t, A.g.T, A.g.X * A.g.Y * A.g.Z, sc, sigma);
        // rows 2,3: my accC (row 2+sc) + partner's accB (its row 3-sc' = 2+sc)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            accC[k].x += __shfl_xor_sync(0xffffffffu, accB[k].x, 16);
            accC[k].y += __shfl_xor_sync(0xffffffffu, accB[k].y, 16);
        }
        const size_t base = (size_t)blk * (12 * 32) + j;
        const double mk = -A.kappa;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int row = half ? 2 + sc : sc;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const size_t idx = base + (size_t)(row * 3 + c) * 32;
                const cplx xi = ldg128(A.in + idx);
                const cplx a = half ? accC[c] : accA[c];
                cplx yk = cmake(fma(mk, a.x, xi.x), fma(mk, a.y, xi.y));
                if (A.fuse.shift_src) {
                    const cplx sv = ldg128(A.fuse.shift_src + idx);
                    yk.x = fma(A.fuse.shift, sv.x, yk.x); yk.y = fma(A.fuse.shift, sv.y, yk.y);
                }
                if (A.fuse.dot_with) {
                    const cplx w = ldg128(A.fuse.dot_with + idx);
                    red[0] = fma(w.x, yk.x, red[0]); red[0] = fma(w.y, yk.y, red[0]);
                    red[1] = fma(w.x, yk.y, red[1]); red[1] = fma(-w.y, yk.x, red[1]);
                }
                red[2] = fma(yk.x, yk.x, red[2]); red[2] = fma(yk.y, yk.y, red[2]);
                A.out[idx] = yk;
            }
        }
    }
    if (A.fuse.dot_with || A.fuse.want_norm) grid_reduce_finish<3>(red, A.red, A.fuse.finish);
}

int launch_wilson_dslash2(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, cudaStream_t s) {
    const int bs = 64 * ctx->g.wpc;                 // two warps per 32-site block
    const int grid = (ctx->g.nblk + ctx->g.wpc - 1) / ctx->g.wpc;
    static int lb = -1;
    if (lb < 0) {
        lb = 0;
        if (const char *e = getenv("LQCD_LB2")) lb = atoi(e);
    }
#define WL2(MT, MB)                                                               \
    do {                                                                          \
        if (dagger) wilson_dslash2_kernel<1, MT, MB><<<grid, bs, 0, s>>>(A);      \
        else        wilson_dslash2_kernel<0, MT, MB><<<grid, bs, 0, s>>>(A);      \
    } while (0)
    if (lb == 2) WL2(256, 2);
    else if (lb == 4) WL2(256, 4);
    else WL2(256, 3);
#undef WL2
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}
