// wilson_eo.cu -- even-odd (Schur complement) preconditioned Wilson solve on checkerboarded half-lattice fields.
//
// New capability behind BASELINE.json configs[1] ("16^4 Wilson Dslash + even-odd CG"); the surveyed wrapper's `isevenodd`
// is a heatbath flag only (src/updates/AbstractUpdate.jl:97), so the algorithm is the textbook one and is
// restated identically in the CPU oracle (oracle/lqcd_oracle.c: orc_eo_solve).  With M = 1 - kappa H, H connecting
// opposite parities:
//     bhat_e = b_e + kappa H_eo b_o;     Mhat x_e = bhat_e,  Mhat = 1 - kappa^2 H_eo H_oe;     x_o = b_o + kappa H_oe x_e
// and |bhat - Mhat x_e|^2 IS the true residual |b - M x|^2 of the full system, so the reference's stopping rule carries over.
//
// Layout: a parity-p half field holds the sites with (x+y+z+t)&1 == p at half index h = (x>>1) + (X/2)*(y + Y*(z + Z*t)),
// AoSoA-32 over h exactly like the full fields (lqcd_internal.cuh); the links are split the same way by the parity of the
// site that owns them.  A hop kernel writes one parity and reads the other: forward links come from the output parity's
// array at the thread's own index (coalesced), backward links from the input parity's array at the neighbour's index.  The
// +-y/z/t neighbours keep their x>>1; the +-x neighbour is h or h+-1 depending on the row parity.
// Bytes per OUTPUT site of one hop: 8 links (1152) + spinor read (192) + write (192) = 1536; Mhat costs two
// hops over V/2 sites each = 1536 B per full-lattice site (M costs 960) and converges in roughly half the iterations on
// vectors of half the length.
//
// Parity on hardware: tests/test_gpu_extended.py (vs orc.eo_solve), tests/test_gpu_baseline_sizes.py (16^4, BASELINE configs[1]);
// index logic also emulated on the CPU (tests/test_evenodd.py::test_checkerboard_index_emulation).
#include "wilson_kernel.cuh"
#include "eo_common.cuh"
#include <cstring>

void make_tiling(Geom &g);                                     // context.cu
int solve_impl(lqcd_ctx *ctx, const lqcd_op *op, cplx *x, const cplx *bb, size_t n, int method, int target,
               double eps, int maxsteps, int *iters, double *resid_sq, double *hist);      // solvers.cu

struct EoArgs {
    cplx *out;               // output half field (parity `parity`)
    const cplx *in;          // input half field (opposite parity)
    const cplx *xsrc;        // nullable: out = xsrc + coef * H in   (same parity as out)
    double coef;
    const cplx *g_out, *g_in;   // links owned by the output / input parity
    Geom gh;
    int parity;
    double bc[4];
    DslashFuse fuse;
    Reduce red;
};

// out(n) = [xsrc(n)] + coef * sum_mu [ (1 -+ g_mu) U_mu(n) in(n+mu) + (1 +- g_mu) U_mu^dag(n-mu) in(n-mu) ],  n of parity `parity`
template <int DAG, int MAXT, int MINB, int LH>
__global__ void __launch_bounds__(MAXT, MINB) wilson_eo_hop_kernel(const EoArgs A) {
    if (A.fuse.use_state && A.red.st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = block_of_warp(A.gh, blockIdx.x, warp);
    const bool active = blk < A.gh.nblk;
    double red[3] = {0.0, 0.0, 0.0};
    if (active) {
        const int h = blk * 32 + lane;
        int xh, y, z, t;
        site_coords(A.gh, h, xh, y, z, t);
        const int Xh = A.gh.X;
        const int odd_row = (y + z + t + A.parity) & 1;          // x = 2*xh + odd_row
        cplx acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = cmake(0.0, 0.0);
        const size_t base = (size_t)blk * (12 * 32) + lane;
        if (A.xsrc || A.fuse.axpy_r || A.fuse.dot_with) {
#pragma unroll
            for (int k = 0; k < 12; k++) {
                if (A.xsrc) prefetch_l2(A.xsrc + base + k * 32);
                if (A.fuse.axpy_r) prefetch_l2(A.fuse.axpy_r + base + k * 32);
                if (A.fuse.dot_with) prefetch_l2(A.fuse.dot_with + base + k * 32);
            }
        }
        {   // x direction: the neighbour of the other parity sits at the same xh or one step away
            const bool wf = odd_row && (xh == Xh - 1);           // x == X-1
            const int nf = odd_row ? (wf ? h - (Xh - 1) : h + 1) : h;
            hop<0, 1, DAG, LH, 0>(acc, A.in, A.g_out, nf, h, wf, A.bc[0]);
            const bool wb = !odd_row && (xh == 0);               // x == 0
            const int nb = odd_row ? h : (wb ? h + (Xh - 1) : h - 1);
            hop<0, 0, DAG, LH, 0>(acc, A.in, A.g_in, nb, nb, wb, A.bc[0]);
        }
        {
            const int st = Xh;
            const bool wf = (y == A.gh.Y - 1), wb = (y == 0);
            const int nf = wf ? h - (A.gh.Y - 1) * st : h + st, nb = wb ? h + (A.gh.Y - 1) * st : h - st;
            hop<1, 1, DAG, LH, 0>(acc, A.in, A.g_out, nf, h, wf, A.bc[1]);
            hop<1, 0, DAG, LH, 0>(acc, A.in, A.g_in, nb, nb, wb, A.bc[1]);
        }
        {
            const int st = Xh * A.gh.Y;
            const bool wf = (z == A.gh.Z - 1), wb = (z == 0);
            const int nf = wf ? h - (A.gh.Z - 1) * st : h + st, nb = wb ? h + (A.gh.Z - 1) * st : h - st;
            hop<2, 1, DAG, LH, 0>(acc, A.in, A.g_out, nf, h, wf, A.bc[2]);
            hop<2, 0, DAG, LH, 0>(acc, A.in, A.g_in, nb, nb, wb, A.bc[2]);
        }
        {
            const int st = Xh * A.gh.Y * A.gh.Z;
            const bool wf = (t == A.gh.T - 1), wb = (t == 0);
            const int nf = wf ? h - (A.gh.T - 1) * st : h + st, nb = wb ? h + (A.gh.T - 1) * st : h - st;
            hop<3, 1, DAG, LH, 0>(acc, A.in, A.g_out, nf, h, wf, A.bc[3]);
            hop<3, 0, DAG, LH, 0>(acc, A.in, A.g_in, nb, nb, wb, A.bc[3]);
        }
        cplx *dst = A.fuse.axpy_r ? A.fuse.axpy_r : A.out;
        const double malpha = A.fuse.axpy_r ? -A.red.st->alpha : 0.0;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            cplx yk = cmake(A.coef * acc[k].x, A.coef * acc[k].y);
            if (A.xsrc) {
                cplx xi = ldg128(A.xsrc + base + k * 32);
                yk = cmake(fma(A.coef, acc[k].x, xi.x), fma(A.coef, acc[k].y, xi.y));
            }
            if (A.fuse.axpy_r) {           // fused CG residual update: r <- r - alpha * y; y itself is not stored
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
    if (A.fuse.dot_with || A.fuse.want_norm) grid_reduce_finish<3>(red, A.red, A.fuse.finish);
}

// ---- layout conversion: full AoSoA-32 <-> the two checkerboard halves ------------------------------------------------
// one thread per (half site, parity); ncomp = 12 (spinor) or 36 (the four links of a site)
template <int TO_HALF>
__global__ void eo_convert_kernel(cplx *full, cplx *he, cplx *ho, Geom g, Geom gh, int ncomp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.V) return;
    const int p = idx / gh.V, h = idx % gh.V;
    int r = h;
    const int xh = r % gh.X; r /= gh.X;
    const int y = r % gh.Y; r /= gh.Y;
    const int z = r % gh.Z;
    const int t = r / gh.Z;
    const int x = 2 * xh + ((y + z + t + p) & 1);
    const int s = x + g.X * (y + g.Y * (z + g.Z * t));
    cplx *F = full + (size_t)(s >> 5) * ncomp * 32 + (s & 31);
    cplx *H = (p ? ho : he) + (size_t)(h >> 5) * ncomp * 32 + (h & 31);
    for (int k = 0; k < ncomp; k++) {
        if (TO_HALF) H[k * 32] = F[k * 32]; else F[k * 32] = H[k * 32];
    }
}

int eo_state(lqcd_ctx *ctx, EoState **out) {
    if (ctx->eo) { *out = ctx->eo; return LQCD_OK; }
    const Geom &g = ctx->g;
    if (ctx->nranks > 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "the even-odd solve is implemented for a single rank");
    if ((g.X | g.Y | g.Z | g.T) & 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "even-odd needs even lattice extents");
    if ((g.V / 2) % 32 != 0) return lqcd_fail(ctx, LQCD_ERR_ARG, "even-odd needs V/2 to be a multiple of 32");
    EoState *e = new EoState();
    memset(e, 0, sizeof *e);
    e->gh = g;
    e->gh.X = g.X / 2; e->gh.gX = g.gX / 2; e->gh.V = g.V / 2; e->gh.nblk = e->gh.V / 32;
    for (int i = 0; i < 4; i++) e->gh.part[i] = 0;
    make_tiling(e->gh);
    e->fullX = g.X;
    e->epoch = ~0ull;
    e->nhalf = (size_t)e->gh.nblk * 12 * 32;
    cudaError_t err = cudaSuccess;
    for (int p = 0; p < 2 && err == cudaSuccess; p++) err = cudaMalloc(&e->gauge[p], (size_t)e->gh.nblk * 36 * 32 * sizeof(cplx));
    for (int i = 0; i < 5 && err == cudaSuccess; i++) err = cudaMalloc(&e->f[i], e->nhalf * sizeof(cplx));
    if (err != cudaSuccess) {
        for (int p = 0; p < 2; p++) cudaFree(e->gauge[p]);
        for (int i = 0; i < 5; i++) cudaFree(e->f[i]);
        delete e;
        return lqcd_fail(ctx, LQCD_ERR_CUDA, "even-odd workspace -> %s", cudaGetErrorString(err));
    }
    ctx->eo = e;
    *out = e;
    return LQCD_OK;
}

void eo_destroy(lqcd_ctx *ctx) {
    EoState *e = ctx->eo;
    if (!e) return;
    for (int p = 0; p < 2; p++) cudaFree(e->gauge[p]);
    for (int i = 0; i < 5; i++) cudaFree(e->f[i]);
    delete e;
    ctx->eo = nullptr;
}

static int eo_hop(lqcd_ctx *ctx, EoState *e, const lqcd_op *op, int dagger, int out_parity, cplx *out, const cplx *in,
                  const cplx *xsrc, double coef, const DslashFuse *fuse) {
    EoArgs A;
    A.out = out; A.in = in; A.xsrc = xsrc; A.coef = coef;
    A.g_out = e->gauge[out_parity]; A.g_in = e->gauge[1 - out_parity];
    A.gh = e->gh; A.parity = out_parity;
    for (int i = 0; i < 4; i++) A.bc[i] = op->bc[i];
    A.fuse = fuse ? *fuse : DslashFuse();
    A.red = ctx->red;
    if (A.fuse.shift_src) return lqcd_fail(ctx, LQCD_ERR_ARG, "even-odd hop: unsupported fused epilogue");
    const int bs = 32 * e->gh.wpc, grid = (e->gh.nblk + e->gh.wpc - 1) / e->gh.wpc;
    if (bs > 256) return lqcd_fail(ctx, LQCD_ERR_ARG, "LQCD_WPC > 8 is not supported");
    const int lh = e->gh.V <= (1 << 17);
#define EK(MT, MB, LH_)                                                                                   \
    do {                                                                                                  \
        if (dagger) wilson_eo_hop_kernel<1, MT, MB, LH_><<<grid, bs, 0, ctx->stream>>>(A);                \
        else        wilson_eo_hop_kernel<0, MT, MB, LH_><<<grid, bs, 0, ctx->stream>>>(A);                \
    } while (0)
    if (bs > 128) { if (lh) EK(256, 1, 1); else EK(256, 1, 0); }
    else          { if (lh) EK(128, 3, 1); else EK(128, 3, 0); }
#undef EK
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}

// y_e = Mhat x_e (dagger: Mhat^dag) with the solver's fused epilogue on the second hop; called by solvers.cu while
// ctx->eo_active is set.  y, x: even half fields.
int eo_mhat(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse) {
    if (op->kind == LQCD_STAGGERED) return stag_even_apply(ctx, op, y, x, dagger, fuse);       // staggered_eo.cu
    EoState *e = ctx->eo;
    DslashFuse plain = DslashFuse();
    plain.use_state = fuse ? fuse->use_state : 0;
    LQCD_TRY(eo_hop(ctx, e, op, dagger, 1, e->f[3], x, nullptr, 1.0, &plain));                 // t_o = H_oe x_e
    return eo_hop(ctx, e, op, dagger, 0, y, e->f[3], x, -op->kappa * op->kappa, fuse);          // y_e = x_e - kappa^2 H_eo t_o
}

int eo_convert(lqcd_ctx *ctx, EoState *e, int to_half, cplx *full, cplx *he, cplx *ho, int ncomp) {
    const int bs = 128, grid = (ctx->g.V + bs - 1) / bs;
    if (to_half) eo_convert_kernel<1><<<grid, bs, 0, ctx->stream>>>(full, he, ho, ctx->g, e->gh, ncomp);
    else         eo_convert_kernel<0><<<grid, bs, 0, ctx->stream>>>(full, he, ho, ctx->g, e->gh, ncomp);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}

extern "C" int lqcd_solve_eo(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *b, int method, int target,
                             double eps, int maxsteps, int *iters, double *resid_sq, double *hist) {
    if (!ctx || !op || !y || !b) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (op->kind != LQCD_WILSON || op->csw != 0.0 || op->r != 1.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "the even-odd solve is built for the Wilson operator with r = 1 and no clover term");
    if (y->owner != ctx || b->owner != ctx || y->kind != LQCD_WILSON || b->kind != LQCD_WILSON || y == b) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad fields");
    if (target != LQCD_OP_D && target != LQCD_OP_DDAG) return lqcd_fail(ctx, LQCD_ERR_ARG, "the even-odd solve handles D x = b or D^dag x = b");
    if (method != LQCD_SOLVER_CGNR && method != LQCD_SOLVER_BICGSTAB) return lqcd_fail(ctx, LQCD_ERR_ARG, "even-odd: method must be CGNR (\"bicg\") or BiCGStab");
    if (maxsteps < 1 || !(eps >= 0.0)) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad eps / maxsteps");
    for (int i = 0; i < 4; i++)
        if (op->bc[i] != 1.0 && op->bc[i] != -1.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "boundary phase bc[%d] = %g must be +-1", i, op->bc[i]);
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "operator applied before lqcd_gauge_upload");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    EoState *e = nullptr;
    LQCD_TRY(eo_state(ctx, &e));
    if (e->epoch != ctx->gauge_epoch) {
        LQCD_TRY(eo_convert(ctx, e, 1, ctx->gauge, e->gauge[0], e->gauge[1], 36));
        e->epoch = ctx->gauge_epoch;
    }
    const int dagger = (target == LQCD_OP_DDAG);
    cplx *be = e->f[0], *bo = e->f[1], *xe = e->f[2], *bh = e->f[4];
    LQCD_TRY(eo_convert(ctx, e, 1, b->d, be, bo, 12));
    LQCD_TRY(eo_convert(ctx, e, 1, y->d, xe, bh, 12));                        // even part of y = initial guess (odd part discarded)
    LQCD_TRY(eo_hop(ctx, e, op, dagger, 0, bh, bo, be, op->kappa, nullptr));    // bhat_e = b_e + kappa H_eo b_o
    ctx->eo_active = 1;
    const int rc = solve_impl(ctx, op, xe, bh, e->nhalf, method, target, eps, maxsteps, iters, resid_sq, hist);
    ctx->eo_active = 0;
    if (rc != LQCD_OK && rc != LQCD_ERR_NOCONV) return rc;
    LQCD_TRY(eo_hop(ctx, e, op, dagger, 1, bh, xe, bo, op->kappa, nullptr));    // x_o = b_o + kappa H_oe x_e  (bhat reused)
    LQCD_TRY(eo_convert(ctx, e, 0, y->d, xe, bh, 12));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}
