// wilson_dslash3.cu -- EXPERIMENTAL (LQCD_WILSON_KERNEL=3, off by default, NOT yet run on hardware):
// t-marching Wilson Dslash with a 3-slice spinor window in shared memory filled by cp.async.bulk (TMA 1-D bulk copies).
//
// Why: the default kernel (wilson_dslash.cu) moves only 1.04x the compulsory bytes from HBM but 2.3x through the L2->SM
// fabric (2.2 KB/site, L1 hit rate 21 %), and that fabric (~11-12 TB/s) is what bounds it (profiles/README.md).  Here a CTA
// owns a small (y,z) patch of 32-site blocks and marches along t: the spinor records of the patch for slices t-1, t, t+1
// live in a shared-memory ring (each 6 KB record arrives by ONE cp.async.bulk with mbarrier completion, issued one step
// ahead), so the +-t neighbours, the site's own spinor and the in-patch +-x/y/z neighbours are served from shared memory
// and each spinor record is read from L2 (Lc+2)/Lc times per chunk instead of ~6.5 times.  Links and out-of-patch
// neighbours still come through the normal load path.  Estimated L2 traffic 1.6 KB/site (-27 %).
//
// The control structure (CTA -> (patch, chunk), warp -> block of the slice, window slot schedule, in-patch test,
// wrap-around and phases) is mirrored line by line by the CPU emulation tests/test_tmarch_emulation.py; names match.
// Single GPU only (no MULTI variants yet).
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "wilson_spin.cuh"
#include "bulk_copy.cuh"
#include <cstdio>
#include <cstdlib>

#define K3_REC (12 * 32)                 // complex numbers per spinor record (one 32-site block)
#define K3_REC_BYTES (K3_REC * 16)

struct K3Args {
    WilsonArgs A;
    int Lc, nchunk, nsb;            // t-steps per CTA, chunks, blocks per t-slice
};

// one hop from explicit pointers: sp -> spinor record component 0 of the neighbour (+lane), component stride 32;
// lk -> link element 0 (+lane), element stride 32.  Pointers may be shared or global (generic loads).
template <int MU, int FWD, int DAG>
__device__ __forceinline__ void hop3(cplx (&acc)[12], const cplx *sp, const cplx *__restrict__ lk, bool wrapped, double phase) {
    constexpr int S = (FWD ^ DAG) ? -1 : +1;
    cplx h0[3], h1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const cplx p0 = sp[(0 + c) * 32], p1 = sp[(3 + c) * 32], p2 = sp[(6 + c) * 32], p3 = sp[(9 + c) * 32];
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
            if (FWD) {
                const cplx u = ldg128(lk + (a * 3 + b) * 32);
                cfma(g0, u, h0[b]); cfma(g1, u, h1[b]);
            } else {
                const cplx u = ldg128(lk + (b * 3 + a) * 32);
                cfmac(g0, u, h0[b]); cfmac(g1, u, h1[b]);
            }
        }
        reconstruct<MU, S>(acc, a, g0, g1);
    }
}

template <int DAG>
__global__ void __launch_bounds__(128, 3) wilson_dslash3_kernel(const K3Args K) {
    const WilsonArgs &A = K.A;
    if (A.fuse.use_state && A.red.st->done) return;
    extern __shared__ __align__(128) unsigned char k3_smem[];
    const Geom &g = A.g;
    const int W = g.c[0] * g.c[1] * g.c[2];                      // warps per CTA = blocks per patch
    cplx *win = reinterpret_cast<cplx *>(k3_smem);               // [3][W][12][32]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(k3_smem + (size_t)3 * W * K3_REC_BYTES);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int npatch = g.nt[0] * g.nt[1] * g.nt[2];
    const int patch = blockIdx.x % npatch, chunk = blockIdx.x / npatch;
    const int p0 = patch % g.nt[0], p1 = (patch / g.nt[0]) % g.nt[1], p2 = patch / (g.nt[0] * g.nt[1]);
    const int w0 = w % g.c[0], w1 = (w / g.c[0]) % g.c[1], w2 = w / (g.c[0] * g.c[1]);
    const int bslice = (p0 * g.c[0] + w0) + g.nb[0] * ((p1 * g.c[1] + w1) + g.nb[1] * (p2 * g.c[2] + w2));
    const int nsb = K.nsb, T = g.T, Lc = K.Lc;
    const int t0 = chunk * Lc;

    if (threadIdx.x == 0) {
        for (int j = 0; j < 3; j++) mbar_init(&mbar[j], (uint32_t)W);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // window loads: rel = slice index relative to t0-1; slot = rel % 3
    auto issue_load = [&](int rel) {
        if (lane == 0) {
            int t = t0 - 1 + rel;
            t = (t + T) % T;
            const int slot = rel % 3;
            mbar_arrive_expect_tx(&mbar[slot], K3_REC_BYTES);
            bulk_g2s(win + ((size_t)slot * W + w) * K3_REC, A.in + ((size_t)bslice + (size_t)t * nsb) * K3_REC, K3_REC_BYTES, &mbar[slot]);
        }
    };
    issue_load(0); issue_load(1); issue_load(2);

    // t-invariant spatial neighbour tables (same formulas as emulate() in tests/test_tmarch_emulation.py)
    const int ssl = bslice * 32 + lane;
    const int x = ssl % g.X, y = (ssl / g.X) % g.Y, z = ssl / (g.X * g.Y);
    int nbl[6], nl[6], nw[6];
    bool wr[6];
    {
        const int coord[3] = {x, y, z}, dim[3] = {g.X, g.Y, g.Z}, stride[3] = {1, g.X, g.X * g.Y};
#pragma unroll
        for (int mu = 0; mu < 3; mu++) {
#pragma unroll
            for (int f = 0; f < 2; f++) {                        // f = 0: forward (+mu), f = 1: backward (-mu)
                const int d = mu * 2 + f;
                const bool wrapd = f == 0 ? (coord[mu] == dim[mu] - 1) : (coord[mu] == 0);
                const int nssl = f == 0 ? (wrapd ? ssl - (dim[mu] - 1) * stride[mu] : ssl + stride[mu])
                                        : (wrapd ? ssl + (dim[mu] - 1) * stride[mu] : ssl - stride[mu]);
                wr[d] = wrapd;
                nbl[d] = nssl >> 5; nl[d] = nssl & 31;
                const int q0 = nbl[d] % g.nb[0], q1 = (nbl[d] / g.nb[0]) % g.nb[1], q2 = nbl[d] / (g.nb[0] * g.nb[1]);
                const bool inp = q0 >= p0 * g.c[0] && q0 < (p0 + 1) * g.c[0] && q1 >= p1 * g.c[1] && q1 < (p1 + 1) * g.c[1] &&
                                 q2 >= p2 * g.c[2] && q2 < (p2 + 1) * g.c[2];
                nw[d] = inp ? (q0 - p0 * g.c[0]) + g.c[0] * ((q1 - p1 * g.c[1]) + g.c[1] * (q2 - p2 * g.c[2])) : -1;
            }
        }
    }

    double red[3] = {0.0, 0.0, 0.0};
    const double mk = -A.kappa;
    const double malpha = A.fuse.axpy_r ? -A.red.st->alpha : 0.0;
    cplx *dst = A.fuse.axpy_r ? A.fuse.axpy_r : A.out;

    for (int r = 1; r <= Lc; r++) {
        const int t = t0 + r - 1;
        // slices t-1, t (first step only) and t+1 must have landed
        if (r == 1) { mbar_wait(&mbar[0], 0); mbar_wait(&mbar[1], 0); }
        mbar_wait(&mbar[(r + 1) % 3], (uint32_t)(((r + 1) / 3) & 1));
        const cplx *cur = win + (size_t)(r % 3) * W * K3_REC;
        const cplx *up = win + (size_t)((r + 1) % 3) * W * K3_REC;
        const cplx *dn = win + (size_t)((r - 1) % 3) * W * K3_REC;
        const size_t blk = (size_t)bslice + (size_t)t * nsb;
        cplx acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = cmake(0.0, 0.0);

#define K3_SPATIAL(MU)                                                                                              \
        {                                                                                                           \
            const int df = MU * 2, db = MU * 2 + 1;                                                                 \
            const cplx *spf = nw[df] >= 0 ? cur + (size_t)nw[df] * K3_REC + nl[df]                                  \
                                          : A.in + ((size_t)nbl[df] + (size_t)t * nsb) * K3_REC + nl[df];           \
            hop3<MU, 1, DAG>(acc, spf, A.gauge + (blk * 4 + MU) * (9 * 32) + lane, wr[df], A.bc[MU]);              \
            const cplx *spb = nw[db] >= 0 ? cur + (size_t)nw[db] * K3_REC + nl[db]                                  \
                                          : A.in + ((size_t)nbl[db] + (size_t)t * nsb) * K3_REC + nl[db];           \
            hop3<MU, 0, DAG>(acc, spb, A.gauge + (((size_t)nbl[db] + (size_t)t * nsb) * 4 + MU) * (9 * 32) + nl[db], \
                             wr[db], A.bc[MU]);                                                                     \
        }
        K3_SPATIAL(0) K3_SPATIAL(1) K3_SPATIAL(2)
#undef K3_SPATIAL
        // t direction: same block position and lane in the neighbouring window slots
        {
            const int tm = (t - 1 + T) % T;
            hop3<3, 1, DAG>(acc, up + (size_t)w * K3_REC + lane, A.gauge + (blk * 4 + 3) * (9 * 32) + lane, t == T - 1, A.bc[3]);
            hop3<3, 0, DAG>(acc, dn + (size_t)w * K3_REC + lane,
                            A.gauge + (((size_t)bslice + (size_t)tm * nsb) * 4 + 3) * (9 * 32) + lane, t == 0, A.bc[3]);
        }
        const size_t base = blk * K3_REC + lane;
        const cplx *own = cur + (size_t)w * K3_REC + lane;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            const cplx xi = own[k * 32];
            cplx yk = cmake(fma(mk, acc[k].x, xi.x), fma(mk, acc[k].y, xi.y));
            if (A.fuse.shift_src) {
                const cplx sv = ldg128(A.fuse.shift_src + base + k * 32);
                yk.x = fma(A.fuse.shift, sv.x, yk.x); yk.y = fma(A.fuse.shift, sv.y, yk.y);
            }
            if (A.fuse.axpy_r) {
                const cplx rv = A.fuse.axpy_r[base + k * 32];
                yk = cmake(fma(malpha, yk.x, rv.x), fma(malpha, yk.y, rv.y));
            }
            if (A.fuse.dot_with) {
                const cplx wv = ldg128(A.fuse.dot_with + base + k * 32);
                red[0] = fma(wv.x, yk.x, red[0]); red[0] = fma(wv.y, yk.y, red[0]);
                red[1] = fma(wv.x, yk.y, red[1]); red[1] = fma(-wv.y, yk.x, red[1]);
            }
            red[2] = fma(yk.x, yk.x, red[2]); red[2] = fma(yk.y, yk.y, red[2]);
            dst[base + k * 32] = yk;
        }
        // everybody is done with slot (r-1)%3 -> refill it with slice r+2 (needed at step r+1 as "t+1")
        __syncthreads();
        if (r + 2 <= Lc + 1) issue_load(r + 2);
    }
    if (A.fuse.dot_with || A.fuse.want_norm) grid_reduce_finish<3>(red, A.red, A.fuse.finish);
}

// returns LQCD_OK if launched; LQCD_ERR_STATE if the geometry does not qualify (caller falls back to kernel 1)
int launch_wilson_dslash3(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, cudaStream_t s) {
    const Geom &g = ctx->g;
    if (!g.regular || g.s[3] != 1 || g.c[3] != 1 || ctx->nranks != 1) return LQCD_ERR_STATE;
    const int W = g.c[0] * g.c[1] * g.c[2];
    if (W * 32 != 128) return LQCD_ERR_STATE;
    const int npatch = g.nt[0] * g.nt[1] * g.nt[2];
    // chunks: enough CTAs for ~2 waves of 3 CTAs/SM, chunk length >= 2, T divisible
    int nchunk = 1;
    while (nchunk * 2 <= g.T / 2 && g.T % (nchunk * 2) == 0 && npatch * nchunk < 2 * 3 * ctx->num_sms) nchunk *= 2;
    if (const char *e = getenv("LQCD_K3_CHUNKS")) { int v = atoi(e); if (v >= 1 && g.T % v == 0 && g.T / v >= 1) nchunk = v; }
    K3Args K;
    K.A = A; K.Lc = g.T / nchunk; K.nchunk = nchunk; K.nsb = g.nb[0] * g.nb[1] * g.nb[2];
    const size_t smem = (size_t)3 * W * K3_REC_BYTES + 3 * sizeof(uint64_t) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(wilson_dslash3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(ctx, cudaFuncSetAttribute(wilson_dslash3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int grid = npatch * nchunk;
    if (dagger) wilson_dslash3_kernel<1><<<grid, 128, smem, s>>>(K);
    else        wilson_dslash3_kernel<0><<<grid, 128, smem, s>>>(K);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}
