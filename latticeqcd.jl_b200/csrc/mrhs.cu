// mrhs.cu -- several right-hand sides per operator application ("multi-RHS"), fp64, sm_100a.
//
// The reference's measurement solves are loops of INDEPENDENT solves against the SAME links: the 12 spin-colour point sources
// of the quark propagator (src/measurements/unusedfiles/measure_Pion_correlator.jl:333-349 maps
// calc_quark_propagators_point_source_each over 1:NC*Nspinor, each ending in solve_DinvX!(p, D, b), :399) and the Nr Z4 noise
// vectors of the chiral condensate (measure_chiral_condensate.jl:176-182).  On the CPU they run one after the other; here they
// run in lock step so that every link matrix fetched from HBM is used for R right-hand sides.  That pays for
// the STAGGERED operator, where links are 576 of the 672 B/site (two-row links: 384 of 480): 576/R + 96 B per site and right-hand
// side.  For Wilson the neighbour spinors dominate the L2 -> SM traffic and sharing links bought nothing (measured, see
// staggered_group below), so Wilson right-hand sides take the single-RHS path inside the same entry points.
//
// Kernel: same site <-> thread, warp <-> AoSoA-32 block and CTA-tile mapping as the single-RHS kernel; per hop the link is loaded
// ONCE into registers and applied to the R neighbour vectors, each RHS keeping its own accumulator.  grid.y runs over groups of R
// right-hand sides.  The per-RHS arithmetic is the single-RHS kernel's, operation for operation, and the per-RHS reductions use the
// same CTA partial order, so one application reproduces lqcd_dslash bit for bit and every right-hand side of a CGNR solve
// reproduces lqcd_solve bit for bit -- iteration counts included (the batched CG on DdagD keeps q = D^dag D p in memory instead
// of fusing the residual update into the second Dslash, so it agrees with the single-RHS CG to rounding).
//
// Solver: CGNR (upstream "bicg", what solve_DinvX!(p, D, b) runs) and CG on DdagD, all right-hand sides advancing together; each
// has its OWN SolverState / reduction workspace in device memory, converged systems drop out of the kernels (mask), the host
// polls all states one batch behind the GPU.  Single rank; anything else takes the single-RHS path, one right-hand side after the other.
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "site_map.cuh"
#include "link_load.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

#define LQCD_MAX_RHS 16

struct MrhsRed { double *partials; unsigned int *ticket; SolverState *st; };

struct MrhsArgs {
    const cplx *in[LQCD_MAX_RHS];
    cplx *out[LQCD_MAX_RHS];
    MrhsRed red[LQCD_MAX_RHS];
    const cplx *gauge;
    const cplx *links12; // two-row links (links12.cu) when the links are SU(3), else null: same choice as the single-RHS kernels
    Geom g;
    double kappa, mass, sign;
    double bc[4];
    int nrhs;
    int want_norm;       // reduce |y_j|^2 (red2) and apply `finish` to right-hand side j's state
    int finish;
    int use_state;       // right-hand sides whose state says done are skipped
};

__device__ __forceinline__ Reduce reduce_of(const MrhsRed &m) {
    Reduce R;
    R.partials = m.partials; R.ticket = m.ticket; R.st = m.st; R.hist = nullptr;
    R.cr.nranks = 0;
    return R;
}

// which of this group's right-hand sides are live (CTA- and grid.y-slice-uniform: `done` is only written by earlier kernels)
template <int R>
__device__ __forceinline__ unsigned live_mask(const MrhsArgs &A, int rhs0) {
    unsigned mask = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int j = rhs0 + r;
        if (j < A.nrhs && !(A.use_state && A.red[j].st->done)) mask |= 1u << r;
    }
    return mask;
}

// ---- staggered -------------------------------------------------------------------------------------------------------------
template <int MU, int FWD, int R>
__device__ __forceinline__ void shop_m(cplx (&acc)[R][3], const MrhsArgs &A, int rhs0, unsigned mask, int ns, int ls, double coef) {
    cplx u[9];
    if (A.links12) load_link<MU, 1, 0>(u, A.links12, ls); else load_link<MU, 0, 0>(u, A.gauge, ls);      // grid-uniform
    const size_t so = (size_t)(ns >> 5) * (3 * 32) + (ns & 31);
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (!((mask >> r) & 1u)) continue;
        const cplx *sp = A.in[rhs0 + r] + so;
        cplx v[3];
#pragma unroll
        for (int c = 0; c < 3; c++) v[c] = cscale(coef, __ldg(sp + c * 32));
#pragma unroll
        for (int a = 0; a < 3; a++) {
#pragma unroll
            for (int b = 0; b < 3; b++) {
                if (FWD) cfma(acc[r][a], u[a * 3 + b], v[b]);
                else     cfmac(acc[r][a], u[b * 3 + a], v[b]);
            }
        }
    }
}

template <int MU, int R>
__device__ __forceinline__ void shop_pair_m(cplx (&acc)[R][3], const MrhsArgs &A, int rhs0, unsigned mask, int s, int coord, int dim, int stride, double eta) {
    {
        const bool w = (coord == dim - 1);
        const int ns = w ? s - (dim - 1) * stride : s + stride;
        shop_m<MU, 1, R>(acc, A, rhs0, mask, ns, s, 0.5 * eta * (w ? A.bc[MU] : 1.0));
    }
    {
        const bool w = (coord == 0);
        const int ns = w ? s + (dim - 1) * stride : s - stride;
        shop_m<MU, 0, R>(acc, A, rhs0, mask, ns, ns, -0.5 * eta * (w ? A.bc[MU] : 1.0));
    }
}

template <int R>
__global__ void __launch_bounds__(256) staggered_mrhs_kernel(const MrhsArgs A) {
    const int rhs0 = blockIdx.y * R;
    const unsigned mask = live_mask<R>(A, rhs0);
    if (!mask) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = block_of_warp(A.g, blockIdx.x, warp);
    const bool active = blk < A.g.nblk;
    cplx acc[R][3];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) acc[r][k] = cmake(0.0, 0.0);
    if (active) {
        const int s = blk * 32 + lane;
        int x, y, z, t;
        site_coords(A.g, s, x, y, z, t);
        const int gx = x + A.g.o[0], gy = y + A.g.o[1], gz = z + A.g.o[2];
        shop_pair_m<0, R>(acc, A, rhs0, mask, s, x, A.g.X, 1, 1.0);
        shop_pair_m<1, R>(acc, A, rhs0, mask, s, y, A.g.Y, A.g.X, (gx & 1) ? -1.0 : 1.0);
        shop_pair_m<2, R>(acc, A, rhs0, mask, s, z, A.g.Z, A.g.X * A.g.Y, ((gx + gy) & 1) ? -1.0 : 1.0);
        shop_pair_m<3, R>(acc, A, rhs0, mask, s, t, A.g.T, A.g.X * A.g.Y * A.g.Z, ((gx + gy + gz) & 1) ? -1.0 : 1.0);
    }
    const size_t base = (size_t)blk * (3 * 32) + lane;
    bool reduced = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (!((mask >> r) & 1u)) continue;
        const int j = rhs0 + r;
        double red[3] = {0.0, 0.0, 0.0};
        if (active) {
            const cplx *xin = A.in[j];
            cplx *dst = A.out[j];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const cplx xi = __ldg(xin + base + k * 32);
                const cplx yk = cmake(fma(A.sign, acc[r][k].x, A.mass * xi.x), fma(A.sign, acc[r][k].y, A.mass * xi.y));
                red[2] = fma(yk.x, yk.x, red[2]); red[2] = fma(yk.y, yk.y, red[2]);
                dst[base + k * 32] = yk;
            }
        }
        if (A.want_norm) {
            if (reduced) __syncthreads();
            grid_reduce_finish<3>(red, reduce_of(A.red[j]), A.finish);
            reduced = true;
        }
    }
}

// ---- batched Krylov updates: blockIdx.y = right-hand side; per-RHS arithmetic and partial order of blas.cu's kernels ----------
#define MB_BS 256
struct MVecs {
    cplx *a[LQCD_MAX_RHS], *b[LQCD_MAX_RHS], *c[LQCD_MAX_RHS], *d[LQCD_MAX_RHS];
    MrhsRed red[LQCD_MAX_RHS];
};

// r = b - q ; p = r (optional) ; red0 = |r|^2           (a = b, b = q, c = r, d = p)
__global__ void __launch_bounds__(MB_BS) km_resid_init(const MVecs V, size_t n, int finish) {
    const int j = blockIdx.y;
    const cplx *__restrict__ b = V.a[j], *__restrict__ q = V.b[j];
    cplx *__restrict__ r = V.c[j], *__restrict__ p = V.d[j];
    double red[3] = {0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx v = csub(b[i], q[i]);
        r[i] = v;
        if (p) p[i] = v;
        red[0] = fma(v.x, v.x, red[0]); red[0] = fma(v.y, v.y, red[0]);
    }
    red[1] = red[0];
    grid_reduce_finish<3>(red, reduce_of(V.red[j]), finish);
}
// CGNR: res -= alpha q ; x += alpha p ; red0 = |res|^2    (a = res, b = q, c = x, d = p)
__global__ void __launch_bounds__(MB_BS) km_nr_update(const MVecs V, size_t n) {
    const int j = blockIdx.y;
    const SolverState *st = V.red[j].st;
    if (st->done) return;
    cplx *__restrict__ res = V.a[j], *__restrict__ x = V.c[j];
    const cplx *__restrict__ q = V.b[j], *__restrict__ p = V.d[j];
    const double alpha = st->alpha;
    double red[1] = {0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx rv = res[i], qv = q[i], xv = x[i], pv = p[i];
        rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        res[i] = rv; x[i] = xv;
        red[0] = fma(rv.x, rv.x, red[0]); red[0] = fma(rv.y, rv.y, red[0]);
    }
    grid_reduce_finish<1>(red, reduce_of(V.red[j]), FIN_NR_RR);
}
// p = beta p + q                                           (a = p, b = q)
__global__ void __launch_bounds__(MB_BS) km_xpby_state(const MVecs V, size_t n) {
    const int j = blockIdx.y;
    const SolverState *st = V.red[j].st;
    if (st->done) return;
    cplx *__restrict__ p = V.a[j];
    const cplx *__restrict__ q = V.b[j];
    const double beta = st->beta;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx pv = p[i], qv = q[i];
        pv.x = fma(beta, pv.x, qv.x); pv.y = fma(beta, pv.y, qv.y);
        p[i] = pv;
    }
}
// CG: r -= alpha q ; red0 = |r|^2                          (a = r, b = q)
__global__ void __launch_bounds__(MB_BS) km_cg_update_r(const MVecs V, size_t n, int finish) {
    const int j = blockIdx.y;
    const SolverState *st = V.red[j].st;
    if (st->done) return;
    cplx *__restrict__ r = V.a[j];
    const cplx *__restrict__ q = V.b[j];
    const double alpha = st->alpha;
    double red[1] = {0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx rv = r[i], qv = q[i];
        rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
        r[i] = rv;
        red[0] = fma(rv.x, rv.x, red[0]); red[0] = fma(rv.y, rv.y, red[0]);
    }
    grid_reduce_finish<1>(red, reduce_of(V.red[j]), finish);
}
// CG: x += alpha p ; p = r + beta p                         (a = x, b = p, c = r); see k_cg_update_xp for the `it` rule
__global__ void __launch_bounds__(MB_BS) km_cg_update_xp(const MVecs V, size_t n, int it) {
    const int j = blockIdx.y;
    const SolverState *st = V.red[j].st;
    const int done = st->done;
    if (done && st->iters < it) return;
    cplx *__restrict__ x = V.a[j], *__restrict__ p = V.b[j];
    const cplx *__restrict__ r = V.c[j];
    const double alpha = st->alpha, beta = st->beta;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx xv = x[i], pv = p[i];
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        x[i] = xv;
        if (!done) {
            cplx rv = r[i];
            pv.x = fma(beta, pv.x, rv.x); pv.y = fma(beta, pv.y, rv.y);
            p[i] = pv;
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
struct MrhsWork {
    SolverState *st_dev;        // [LQCD_MAX_RHS]
    SolverState *st_host;       // pinned, [3][LQCD_MAX_RHS]: init image, two polling slots
    double *partials;           // [LQCD_MAX_RHS][maxgrid * LQCD_MAX_RED]
    unsigned int *tickets;      // [LQCD_MAX_RHS]
    size_t maxgrid;
};

void mrhs_destroy(lqcd_ctx *ctx) {
    MrhsWork *w = ctx->mrhs;
    if (!w) return;
    cudaFree(w->st_dev); cudaFree(w->partials); cudaFree(w->tickets);
    cudaFreeHost(w->st_host);
    delete w;
    ctx->mrhs = nullptr;
}

static int mrhs_work(lqcd_ctx *ctx, MrhsWork **out) {
    if (!ctx->mrhs) {
        MrhsWork *w = new MrhsWork();
        memset(w, 0, sizeof *w);
        ctx->mrhs = w;                      // owned by the context from here on (freed by mrhs_destroy even if an allocation fails)
        w->maxgrid = (size_t)ctx->g.nblk + 1024;
        CUDA_TRY(ctx, cudaMalloc(&w->st_dev, LQCD_MAX_RHS * sizeof(SolverState)));
        CUDA_TRY(ctx, cudaMemset(w->st_dev, 0, LQCD_MAX_RHS * sizeof(SolverState)));
        CUDA_TRY(ctx, cudaMalloc(&w->partials, LQCD_MAX_RHS * w->maxgrid * LQCD_MAX_RED * sizeof(double)));
        CUDA_TRY(ctx, cudaMalloc(&w->tickets, LQCD_MAX_RHS * sizeof(unsigned int)));
        CUDA_TRY(ctx, cudaMemset(w->tickets, 0, LQCD_MAX_RHS * sizeof(unsigned int)));
        CUDA_TRY(ctx, cudaMallocHost(&w->st_host, 3 * LQCD_MAX_RHS * sizeof(SolverState)));
    }
    if (!ctx->mrhs->st_dev || !ctx->mrhs->partials || !ctx->mrhs->tickets || !ctx->mrhs->st_host)
        return lqcd_fail(ctx, LQCD_ERR_STATE, "multi-RHS workspace was not allocated (earlier CUDA error)");
    *out = ctx->mrhs;
    return LQCD_OK;
}

static void fill_red(const MrhsWork *w, MrhsRed *red) {
    for (int j = 0; j < LQCD_MAX_RHS; j++) {
        red[j].partials = w->partials + (size_t)j * w->maxgrid * LQCD_MAX_RED;
        red[j].ticket = w->tickets + j;
        red[j].st = w->st_dev + j;
    }
}

// Right-hand sides per thread of the staggered kernel: 2 / 3 / 4 / 6 / 8 / 12 (LQCD_MRHS_R_STAGGERED), default 4.  Measured on B200 at
// 32^4 with 12 right-hand sides and two-row links (round 2, profiles/r2e_experiments_n1.json), per right-hand side against the
// single-RHS kernel: R = 2 1.08x, 3 1.23x, 4 1.34x, 6 1.08x, 12 0.81x (230 registers).
// A Wilson multi-RHS kernel (links in registers shared by R = 2 / 3 / 4 right-hand sides, and a variant with the links staged in
// shared memory by bulk copies) was measured in rounds 1 and 2: 0.56-1.04x -- the Wilson kernel is bound by spinor traffic over the
// L2 -> SM fabric, which sharing links does not reduce -- and removed; Wilson right-hand sides go through the single-RHS path.
static int staggered_group(int nrhs) {
    static int env = -1;
    if (env < 0) { const char *e = getenv("LQCD_MRHS_R_STAGGERED"); env = e ? atoi(e) : 0; }
    if (env == 2 || env == 3 || env == 4 || env == 6 || env == 8 || env == 12) return env;
    return nrhs <= 2 ? 2 : (nrhs == 3 ? 3 : 4);          // measured at 32^4, 12 right-hand sides (B200, round 2): R = 4 1.39x, 6 1.12x, 12 0.81x per RHS
}

// one Dslash of all right-hand sides: out[j] = D in[j] (dagger: D^dag); want_norm -> |out[j]|^2 reduced, `finish` applied per RHS
static int launch_mrhs(lqcd_ctx *ctx, const lqcd_op *op, cplx *const *out, const cplx *const *in, int nrhs, int dagger,
                       int want_norm, int finish, int use_state) {
    MrhsWork *w = nullptr;
    LQCD_TRY(mrhs_work(ctx, &w));
    MrhsArgs A;
    memset(&A, 0, sizeof A);
    for (int j = 0; j < nrhs; j++) {
        if (in[j] == out[j]) return lqcd_fail(ctx, LQCD_ERR_ARG, "dslash: in-place application is not allowed");
        A.in[j] = in[j]; A.out[j] = out[j];
    }
    fill_red(w, A.red);
    int g12 = 0;
    LQCD_TRY(ensure_links12(ctx, &g12));
    A.links12 = g12 ? ctx->links12 : nullptr;
    A.gauge = ctx->gauge; A.g = ctx->g; A.kappa = op->kappa; A.mass = op->mass; A.sign = dagger ? -1.0 : 1.0;
    for (int i = 0; i < 4; i++) A.bc[i] = op->bc[i];
    A.nrhs = nrhs; A.want_norm = want_norm; A.finish = finish; A.use_state = use_state;
    const int bs = 32 * ctx->g.wpc;
    const int gx = (ctx->g.nblk + ctx->g.wpc - 1) / ctx->g.wpc;
    if (op->kind != LQCD_STAGGERED) return lqcd_fail(ctx, LQCD_ERR_STATE, "multi-RHS kernel: staggered only");
    {
        const int R = staggered_group(nrhs);
        const dim3 grid(gx, (nrhs + R - 1) / R);
        if (R == 2) staggered_mrhs_kernel<2><<<grid, bs, 0, ctx->stream>>>(A);
        else if (R == 3) staggered_mrhs_kernel<3><<<grid, bs, 0, ctx->stream>>>(A);
        else if (R == 4) staggered_mrhs_kernel<4><<<grid, bs, 0, ctx->stream>>>(A);
        else if (R == 6) staggered_mrhs_kernel<6><<<grid, bs, 0, ctx->stream>>>(A);
        else if (R == 8) staggered_mrhs_kernel<8><<<grid, bs, 0, ctx->stream>>>(A);
        else staggered_mrhs_kernel<12><<<grid, bs, 0, ctx->stream>>>(A);
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}

static inline int mb_grid(const lqcd_ctx *ctx, size_t n) {      // = blas_grid (blas.cu): same partial order per right-hand side
    size_t need = (n + MB_BS - 1) / MB_BS, cap = (size_t)ctx->num_sms * 8;
    return (int)(need < cap ? need : cap);
}
#define MLAUNCH(kernel, nrhs, n, ...)                                                   \
    do {                                                                                \
        kernel<<<dim3(mb_grid(ctx, n), nrhs), MB_BS, 0, ctx->stream>>>(__VA_ARGS__);    \
        ctx->launches++;                                                                \
        CUDA_TRY(ctx, cudaGetLastError());                                              \
    } while (0)

static int check_fields(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *const ys[], const lqcd_fermion *const xs[], int nrhs) {
    if (!ctx || !op || !ys || !xs) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (nrhs < 1 || nrhs > LQCD_MAX_RHS) return lqcd_fail(ctx, LQCD_ERR_ARG, "nrhs must be in [1, %d]", LQCD_MAX_RHS);
    if (op->kind != LQCD_WILSON && op->kind != LQCD_STAGGERED) return lqcd_fail(ctx, LQCD_ERR_ARG, "unknown operator kind %d", op->kind);
    for (int i = 0; i < 4; i++)
        if (op->bc[i] != 1.0 && op->bc[i] != -1.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "boundary phase bc[%d] = %g must be +-1", i, op->bc[i]);
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "operator applied before lqcd_gauge_upload");
    for (int j = 0; j < nrhs; j++) {
        if (!ys[j] || !xs[j]) return lqcd_fail(ctx, LQCD_ERR_ARG, "null field %d", j);
        if (ys[j]->owner != ctx || xs[j]->owner != ctx) return lqcd_fail(ctx, LQCD_ERR_ARG, "field %d belongs to another context", j);
        if (ys[j]->kind != op->kind || xs[j]->kind != op->kind) return lqcd_fail(ctx, LQCD_ERR_ARG, "fermion kind of field %d does not match the operator", j);
        if (ys[j] == xs[j]) return lqcd_fail(ctx, LQCD_ERR_ARG, "output %d aliases its input", j);
        for (int k = 0; k < j; k++)
            if (ys[j] == ys[k] || ys[j] == xs[k] || xs[j] == ys[k]) return lqcd_fail(ctx, LQCD_ERR_ARG, "fields %d and %d alias", j, k);
    }
    return LQCD_OK;
}

// the batched kernels cover: one rank, staggered, full (not even-odd) fields, regular or irregular tiling
static bool batched_ok(const lqcd_ctx *ctx, const lqcd_op *op) {
    return ctx->nranks == 1 && op->kind == LQCD_STAGGERED && !ctx->eo_active && 32 * ctx->g.wpc <= 256;
}

extern "C" int lqcd_dslash_multi(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *const ys[], const lqcd_fermion *const xs[], int nrhs, int mode) {
    LQCD_TRY(check_fields(ctx, op, ys, xs, nrhs));
    if (mode < 0 || mode > 2) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad mode %d", mode);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!batched_ok(ctx, op)) {
        for (int j = 0; j < nrhs; j++) LQCD_TRY(lqcd_dslash(ctx, op, ys[j], xs[j], mode));
        return LQCD_OK;
    }
    cplx *out[LQCD_MAX_RHS];
    const cplx *in[LQCD_MAX_RHS];
    for (int j = 0; j < nrhs; j++) { out[j] = ys[j]->d; in[j] = xs[j]->d; }
    if (mode == LQCD_OP_DDAGD) {
        cplx *tmp[LQCD_MAX_RHS];
        for (int j = 0; j < nrhs; j++) {
            lqcd_fermion *t = nullptr;
            LQCD_TRY(get_scratch(ctx, op->kind, SCR_MRHS0 + j, &t));
            tmp[j] = t->d;
        }
        LQCD_TRY(launch_mrhs(ctx, op, tmp, in, nrhs, 0, 0, 0, 0));
        LQCD_TRY(launch_mrhs(ctx, op, out, tmp, nrhs, 1, 0, 0, 0));
    } else {
        LQCD_TRY(launch_mrhs(ctx, op, out, in, nrhs, mode == LQCD_OP_DDAG, 0, 0, 0));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

// ---- lock-step Krylov loops --------------------------------------------------------------------------------------------------
// Polls all right-hand sides' states one batch behind the enqueue front (same scheme as run_loop in solvers.cu).
template <class Body>
static int run_loop_multi(lqcd_ctx *ctx, MrhsWork *w, int nrhs, int maxsteps, Body body, int *iters, double *resid_sq) {
    int batch = 8;
    if (const char *e = getenv("LQCD_CG_BATCH")) { int v = atoi(e); if (v >= 1 && v <= 1024) batch = v; }
    cudaEvent_t ev[2] = {ctx->ev_poll[0], ctx->ev_poll[1]};
    const size_t snap = (size_t)nrhs * sizeof(SolverState);
    auto slot = [&](int k) { return w->st_host + (size_t)k * LQCD_MAX_RHS; };
    auto all_done = [&](const SolverState *s) { for (int j = 0; j < nrhs; j++) if (!s[j].done) return false; return true; };
    CUDA_TRY(ctx, cudaMemcpyAsync(slot(1), w->st_dev, snap, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ev[0], ctx->stream));
    int it = 0, pending = 1;
    const SolverState *fin = slot(1);
    bool done = false;
    while (true) {
        const int prev = pending;
        bool enq = false;
        if (it < maxsteps) {
            const int hi = it + batch < maxsteps ? it + batch : maxsteps;
            for (int i = it + 1; i <= hi; i++) LQCD_TRY(body(i));
            it = hi; enq = true;
            pending = (pending == 1) ? 2 : 1;
            CUDA_TRY(ctx, cudaMemcpyAsync(slot(pending), w->st_dev, snap, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaEventRecord(ev[pending - 1], ctx->stream));
        }
        CUDA_TRY(ctx, cudaEventSynchronize(ev[prev - 1]));
        fin = slot(prev);
        if (all_done(fin)) { done = true; break; }
        if (!enq) break;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (!done) { fin = slot(pending); done = all_done(fin); }
    int bad = -1, broke = -1;
    for (int j = 0; j < nrhs; j++) {
        if (iters) iters[j] = fin[j].done ? fin[j].iters : fin[j].it;
        if (resid_sq) resid_sq[j] = fin[j].rr;
        if (!fin[j].done && bad < 0) bad = j;
        if (fin[j].done && fin[j].failed && broke < 0) broke = j;
    }
    if (broke >= 0) return lqcd_fail(ctx, LQCD_ERR_NOCONV, "Krylov breakdown on right-hand side %d: |r|^2 is not finite at step %d", broke, fin[broke].it);
    if (bad >= 0) return lqcd_fail(ctx, LQCD_ERR_NOCONV, "solver not converged after %d steps on right-hand side %d (|r|^2 = %.6e, eps = %.3e)",
                                   fin[bad].it, bad, fin[bad].rr, fin[bad].eps);
    return LQCD_OK;
}

extern "C" int lqcd_solve_multi(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *const ys[], const lqcd_fermion *const bs[], int nrhs,
                                int method, int target, double eps, int maxsteps, int *iters, double *resid_sq) {
    LQCD_TRY(check_fields(ctx, op, ys, bs, nrhs));
    if (maxsteps < 1 || !(eps >= 0.0)) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad eps / maxsteps");
    if (method == LQCD_SOLVER_CG && target != LQCD_OP_DDAGD) return lqcd_fail(ctx, LQCD_ERR_ARG, "CG needs the Hermitian target DdagD");
    if (method != LQCD_SOLVER_CG && target != LQCD_OP_D && target != LQCD_OP_DDAG) return lqcd_fail(ctx, LQCD_ERR_ARG, "CGNR/BiCGStab solve D or D^dag");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!batched_ok(ctx, op) || method == LQCD_SOLVER_BICGSTAB) {       // one right-hand side after the other through the single-RHS path
        int first_bad = -1;
        std::string first_msg;
        for (int j = 0; j < nrhs; j++) {
            int it = 0; double rs = 0.0;
            const int rc = lqcd_solve(ctx, op, ys[j], bs[j], method, target, eps, maxsteps, &it, &rs, nullptr);
            if (iters) iters[j] = it;
            if (resid_sq) resid_sq[j] = rs;
            if (rc != LQCD_OK && rc != LQCD_ERR_NOCONV) return rc;
            if (rc != LQCD_OK && first_bad < 0) { first_bad = j; first_msg = ctx->err; }      // keep solving the others, like the batched path
        }
        if (first_bad >= 0) return lqcd_fail(ctx, LQCD_ERR_NOCONV, "right-hand side %d: %s", first_bad, first_msg.c_str());
        return LQCD_OK;
    }
    MrhsWork *w = nullptr;
    LQCD_TRY(mrhs_work(ctx, &w));
    const int kind = op->kind;
    const size_t n = (size_t)ctx->g.nblk * ys[0]->ncomp * 32;
    // per-RHS work vectors: scratch slots SCR_MRHS0 + 16 v + j (v = 0 .. 3), allocated on first use
    cplx *x[LQCD_MAX_RHS], *v0[LQCD_MAX_RHS], *v1[LQCD_MAX_RHS], *v2[LQCD_MAX_RHS], *v3[LQCD_MAX_RHS];
    const cplx *b[LQCD_MAX_RHS];
    const int nvec = method == LQCD_SOLVER_CG ? 4 : 3;
    for (int j = 0; j < nrhs; j++) {
        x[j] = ys[j]->d; b[j] = bs[j]->d;
        cplx **dst[4] = {&v0[j], &v1[j], &v2[j], &v3[j]};
        for (int v = 0; v < nvec; v++) {
            lqcd_fermion *f = nullptr;
            LQCD_TRY(get_scratch(ctx, kind, SCR_MRHS0 + 16 * v + j, &f));
            *dst[v] = f->d;
        }
    }
    // states
    SolverState *h = w->st_host;
    for (int j = 0; j < nrhs; j++) {
        memset(&h[j], 0, sizeof h[j]);
        h[j].eps = eps; h[j].maxit = maxsteps; h[j].alpha_old = 1.0;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(w->st_dev, h, (size_t)nrhs * sizeof(SolverState), cudaMemcpyHostToDevice, ctx->stream));
    MVecs V;
    memset(&V, 0, sizeof V);
    fill_red(w, V.red);
    auto vec = [&](cplx *const *a, cplx *const *bb, cplx *const *c, cplx *const *d) {
        for (int j = 0; j < nrhs; j++) { V.a[j] = a ? a[j] : nullptr; V.b[j] = bb ? bb[j] : nullptr; V.c[j] = c ? c[j] : nullptr; V.d[j] = d ? d[j] : nullptr; }
    };
    cplx *const *bq = const_cast<cplx *const *>(b);
    if (method == LQCD_SOLVER_CGNR) {
        const int dagA = target == LQCD_OP_DDAG, dagAd = !dagA;
        cplx **res = v0, **p = v1, **q = v2;
        LQCD_TRY(launch_mrhs(ctx, op, q, x, nrhs, dagA, 0, 0, 0));                                  // q = A x0
        vec(bq, q, res, nullptr);
        MLAUNCH(km_resid_init, nrhs, n, V, n, FIN_CG_INIT);                                         // res = b - q, |res|^2
        LQCD_TRY(launch_mrhs(ctx, op, q, res, nrhs, dagAd, 1, FIN_NR_C1, 1));                       // q = A^dag res, c1 = |q|^2
        for (int j = 0; j < nrhs; j++) CUDA_TRY(ctx, cudaMemcpyAsync(p[j], q[j], n * sizeof(cplx), cudaMemcpyDeviceToDevice, ctx->stream));
        return run_loop_multi(ctx, w, nrhs, maxsteps, [&](int) -> int {
            LQCD_TRY(launch_mrhs(ctx, op, q, p, nrhs, dagA, 1, FIN_NR_C2, 1));                      // q = A p, alpha = c1/|q|^2
            vec(res, q, x, p);
            MLAUNCH(km_nr_update, nrhs, n, V, n);                                                   // res -= alpha q, x += alpha p, |res|^2
            LQCD_TRY(launch_mrhs(ctx, op, q, res, nrhs, dagAd, 1, FIN_NR_C3, 1));                   // q = A^dag res, beta = |q|^2/c1
            vec(p, q, nullptr, nullptr);
            MLAUNCH(km_xpby_state, nrhs, n, V, n);                                                  // p = beta p + q
            return LQCD_OK;
        }, iters, resid_sq);
    }
    // CG on DdagD, 4 kernels per iteration: t = D p (|t|^2 = <p, DdagD p> -> alpha), q = D^dag t, r -= alpha q (|r|^2, beta), x/p update
    cplx **r = v0, **p = v1, **q = v2, **t = v3;
    LQCD_TRY(launch_mrhs(ctx, op, t, x, nrhs, 0, 0, 0, 0));
    LQCD_TRY(launch_mrhs(ctx, op, q, t, nrhs, 1, 0, 0, 0));                                         // q = DdagD x0
    vec(bq, q, r, p);
    MLAUNCH(km_resid_init, nrhs, n, V, n, FIN_CG_INIT);                                             // r = p = b - q
    return run_loop_multi(ctx, w, nrhs, maxsteps, [&](int it) -> int {
        LQCD_TRY(launch_mrhs(ctx, op, t, p, nrhs, 0, 1, FIN_CG_PQN, 1));
        LQCD_TRY(launch_mrhs(ctx, op, q, t, nrhs, 1, 0, 0, 1));
        vec(r, q, nullptr, nullptr);
        MLAUNCH(km_cg_update_r, nrhs, n, V, n, FIN_CG_RR);
        vec(x, p, r, nullptr);
        MLAUNCH(km_cg_update_xp, nrhs, n, V, n, it);
        return LQCD_OK;
    }, iters, resid_sq);
}
