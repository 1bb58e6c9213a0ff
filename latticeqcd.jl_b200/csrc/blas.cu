// blas.cu -- BLAS-1 on fermion fields and the fused vector updates of the Krylov loops.
//
// Replaces LatticeDiracOperators.jl's add!/axpby/dot/clear_fermion! on pseudofermion fields
// (reference call sites src/updates/standardHMC.jl:54 dot(xi,xi),
// src/md/standardMD.jl:50-51 similar(), measure_Pion_correlator.jl:370-371 clear_fermion!).
// The layout is irrelevant for element-wise work: fields are flat arrays of double2, streamed with
// 128-bit loads/stores by a grid sized to the SM count (grid-stride, 4 independent loads in flight per
// thread).  Scalars (alpha, beta, ...) are read from SolverState in device memory so the solver loops
// never wait for the host.
#include "lqcd_internal.cuh"
#include "reduce.cuh"

#define BLAS_BS 256

static inline int blas_grid(const lqcd_ctx *ctx, size_t n) {
    size_t need = (n + BLAS_BS - 1) / BLAS_BS;
    size_t cap = (size_t)ctx->num_sms * 8;
    return (int)(need < cap ? need : cap);
}

__global__ void __launch_bounds__(BLAS_BS) k_axpy(cplx a, const cplx *__restrict__ x, cplx *__restrict__ y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx xv = x[i], yv = y[i];
        cfma(yv, a, xv);
        y[i] = yv;
    }
}
__global__ void __launch_bounds__(BLAS_BS) k_xpby(const cplx *__restrict__ x, cplx b, cplx *__restrict__ y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx xv = x[i], yv = y[i];
        cfma(xv, b, yv);
        y[i] = xv;
    }
}
__global__ void __launch_bounds__(BLAS_BS) k_scale(cplx a, cplx *__restrict__ x, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = cmul(a, x[i]);
}
// red0/1 = <a,b>, red2 = |a|^2
__global__ void __launch_bounds__(BLAS_BS) k_dot(const cplx *__restrict__ a, const cplx *__restrict__ b, size_t n, Reduce R, int finish) {
    double red[3] = {0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx av = a[i], bv = b[i];
        red[0] = fma(av.x, bv.x, red[0]); red[0] = fma(av.y, bv.y, red[0]);
        red[1] = fma(av.x, bv.y, red[1]); red[1] = fma(-av.y, bv.x, red[1]);
        red[2] = fma(av.x, av.x, red[2]); red[2] = fma(av.y, av.y, red[2]);
    }
    grid_reduce_finish<3>(red, R, finish);
}

// red0/1 = <w,y> (w may be null), red2 = |y|^2 -- the multi-rank Dslash epilogue (after the halo update)
__global__ void __launch_bounds__(BLAS_BS) k_dot2(const cplx *__restrict__ w, const cplx *__restrict__ y, size_t n, Reduce R, int finish, int use_state) {
    if (use_state && R.st->done) return;
    double red[3] = {0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx yv = y[i];
        if (w) {
            cplx wv = w[i];
            red[0] = fma(wv.x, yv.x, red[0]); red[0] = fma(wv.y, yv.y, red[0]);
            red[1] = fma(wv.x, yv.y, red[1]); red[1] = fma(-wv.y, yv.x, red[1]);
        }
        red[2] = fma(yv.x, yv.x, red[2]); red[2] = fma(yv.y, yv.y, red[2]);
    }
    grid_reduce_finish<3>(red, R, finish);
}

// ---- solver kernels -----------------------------------------------------------------------------------
// r = b - q ; p = r (optional) ; r0 = r (optional) ; red0 = |r|^2 ; red1/2 = <r0,r> = (|r|^2, 0)
__global__ void __launch_bounds__(BLAS_BS) k_resid_init(const cplx *__restrict__ b, const cplx *__restrict__ q,
                                                        cplx *__restrict__ r, cplx *__restrict__ p, cplx *__restrict__ r0,
                                                        size_t n, Reduce R, int finish) {
    double red[3] = {0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx v = csub(b[i], q[i]);
        r[i] = v;
        if (p) p[i] = v;
        if (r0) r0[i] = v;
        red[0] = fma(v.x, v.x, red[0]); red[0] = fma(v.y, v.y, red[0]);
    }
    red[1] = red[0];
    grid_reduce_finish<3>(red, R, finish);
}

// CG: r -= alpha q ; red0 = |r|^2                                   (FIN_CG_RR / FIN_MS_RR)
// multishift additionally: x_j += alpha_j p_j for all shifts (xs/ps arrays of pointers in device memory)
__global__ void __launch_bounds__(BLAS_BS) k_cg_update_r(cplx *__restrict__ r, const cplx *__restrict__ q, size_t n, Reduce R, int finish) {
    const SolverState *st = R.st;
    if (st->done) return;
    const double alpha = st->alpha;
    double red[1] = {0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx rv = r[i], qv = q[i];
        rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
        r[i] = rv;
        red[0] = fma(rv.x, rv.x, red[0]); red[0] = fma(rv.y, rv.y, red[0]);
    }
    grid_reduce_finish<1>(red, R, finish);
}
// CG: x += alpha p ; p = r + beta p.   `it` = iteration this launch belongs to: on the converging
// iteration only the x update is applied (the reference updates x before testing |r|^2, App. C.3).
__global__ void __launch_bounds__(BLAS_BS) k_cg_update_xp(cplx *__restrict__ x, cplx *__restrict__ p, const cplx *__restrict__ r,
                                                          size_t n, const SolverState *__restrict__ st, int it) {
    const int done = st->done;
    if (done && st->iters < it) return;
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
// CGNR: res -= alpha q ; x += alpha p ; red0 = |res|^2                (FIN_NR_RR)
__global__ void __launch_bounds__(BLAS_BS) k_nr_update(cplx *__restrict__ res, const cplx *__restrict__ q, cplx *__restrict__ x,
                                                       const cplx *__restrict__ p, size_t n, Reduce R) {
    const SolverState *st = R.st;
    if (st->done) return;
    const double alpha = st->alpha;
    double red[1] = {0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx rv = res[i], qv = q[i], xv = x[i], pv = p[i];
        rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        res[i] = rv; x[i] = xv;
        red[0] = fma(rv.x, rv.x, red[0]); red[0] = fma(rv.y, rv.y, red[0]);
    }
    grid_reduce_finish<1>(red, R, FIN_NR_RR);
}
// p = beta p + q  (beta from state)
__global__ void __launch_bounds__(BLAS_BS) k_xpby_state(cplx *__restrict__ p, const cplx *__restrict__ q, size_t n, const SolverState *__restrict__ st) {
    if (st->done) return;
    const double beta = st->beta;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx pv = p[i], qv = q[i];
        pv.x = fma(beta, pv.x, qv.x); pv.y = fma(beta, pv.y, qv.y);
        p[i] = pv;
    }
}
// BiCGStab: s = r - alpha v
__global__ void __launch_bounds__(BLAS_BS) k_bi_s(cplx *__restrict__ s, const cplx *__restrict__ r, const cplx *__restrict__ v, size_t n, const SolverState *__restrict__ st) {
    if (st->done) return;
    const cplx ma = make_double2(-st->alpha_re, -st->alpha_im);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx sv = r[i];
        cfma(sv, ma, v[i]);
        s[i] = sv;
    }
}
// BiCGStab: x += alpha p + omega s ; r = s - omega t ; red0 = |r|^2, red1/2 = <r0, r>     (FIN_BI_RR)
__global__ void __launch_bounds__(BLAS_BS) k_bi_xr(cplx *__restrict__ x, const cplx *__restrict__ p, const cplx *__restrict__ s,
                                                   cplx *__restrict__ r, const cplx *__restrict__ t, const cplx *__restrict__ r0,
                                                   size_t n, Reduce R) {
    const SolverState *st = R.st;
    if (st->done) return;
    const cplx al = make_double2(st->alpha_re, st->alpha_im), om = make_double2(st->omega_re, st->omega_im);
    const cplx mom = make_double2(-om.x, -om.y);
    double red[3] = {0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx xv = x[i], sv = s[i];
        cfma(xv, al, p[i]); cfma(xv, om, sv);
        x[i] = xv;
        cplx rv = sv;
        cfma(rv, mom, t[i]);
        r[i] = rv;
        cplx r0v = r0[i];
        red[0] = fma(rv.x, rv.x, red[0]); red[0] = fma(rv.y, rv.y, red[0]);
        red[1] = fma(r0v.x, rv.x, red[1]); red[1] = fma(r0v.y, rv.y, red[1]);
        red[2] = fma(r0v.x, rv.y, red[2]); red[2] = fma(-r0v.y, rv.x, red[2]);
    }
    grid_reduce_finish<3>(red, R, FIN_BI_RR);
}
// BiCGStab: p = r + beta (p - omega v)
__global__ void __launch_bounds__(BLAS_BS) k_bi_p(cplx *__restrict__ p, const cplx *__restrict__ r, const cplx *__restrict__ v, size_t n, const SolverState *__restrict__ st) {
    if (st->done) return;
    const cplx be = make_double2(st->bre, st->bim), mom = make_double2(-st->omega_re, -st->omega_im);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx pv = p[i];
        cfma(pv, mom, v[i]);
        cplx out = r[i];
        cfma(out, be, pv);
        p[i] = out;
    }
}
// multi-shift: x_j += alpha_j p_j ; p_j = zeta_j r + beta_j p_j  (all shifts in one pass over r).
// On the converging iteration only the x update is applied (as in the reference's shiftedcg order).
__global__ void __launch_bounds__(BLAS_BS) k_ms_update_xp(MSPtrs P, const cplx *__restrict__ r, size_t n, const SolverState *__restrict__ st, int it) {
    const int done = st->done;
    if (done && st->iters < it) return;
    const int ns = st->nshift;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx rv = r[i];
        for (int j = 0; j < ns; j++) {
            const double a = st->alpha_s[j];
            cplx xv = P.x[j][i], pv = P.p[j][i];
            xv.x = fma(a, pv.x, xv.x); xv.y = fma(a, pv.y, xv.y);
            P.x[j][i] = xv;
            if (!done) {
                const double z = (j == 0) ? 1.0 : st->zeta[j], b = st->beta_s[j];
                pv.x = fma(b, pv.x, z * rv.x); pv.y = fma(b, pv.y, z * rv.y);
                P.p[j][i] = pv;
            }
        }
    }
}

// ---- host wrappers used by solvers.cu ------------------------------------------------------------------
#define LAUNCH(ctx, kernel, n, ...)                                                                      \
    do {                                                                                                 \
        kernel<<<blas_grid(ctx, n), BLAS_BS, 0, (ctx)->stream>>>(__VA_ARGS__);                           \
        (ctx)->launches++;                                                                               \
        CUDA_TRY(ctx, cudaGetLastError());                                                               \
    } while (0)

int blas_zero(lqcd_ctx *ctx, cplx *x, size_t n) {
    CUDA_TRY(ctx, cudaMemsetAsync(x, 0, n * sizeof(cplx), ctx->stream));
    return LQCD_OK;
}
int blas_copy(lqcd_ctx *ctx, cplx *dst, const cplx *src, size_t n) {
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, n * sizeof(cplx), cudaMemcpyDeviceToDevice, ctx->stream));
    return LQCD_OK;
}
int blas_dot_async(lqcd_ctx *ctx, const cplx *a, const cplx *b, size_t n, int finish) {
    LAUNCH(ctx, k_dot, n, a, b, n, ctx->red, finish);
    return LQCD_OK;
}
int blas_dot2_async(lqcd_ctx *ctx, const cplx *w, const cplx *y, size_t n, int finish, int use_state) {
    LAUNCH(ctx, k_dot2, n, w, y, n, ctx->red, finish, use_state);
    return LQCD_OK;
}
int blas_resid_init(lqcd_ctx *ctx, const cplx *b, const cplx *q, cplx *r, cplx *p, cplx *r0, size_t n, int finish) {
    LAUNCH(ctx, k_resid_init, n, b, q, r, p, r0, n, ctx->red, finish);
    return LQCD_OK;
}
int blas_cg_update_r(lqcd_ctx *ctx, cplx *r, const cplx *q, size_t n, int finish) {
    LAUNCH(ctx, k_cg_update_r, n, r, q, n, ctx->red, finish);
    return LQCD_OK;
}
int blas_cg_update_xp(lqcd_ctx *ctx, cplx *x, cplx *p, const cplx *r, size_t n, int it) {
    LAUNCH(ctx, k_cg_update_xp, n, x, p, r, n, ctx->red.st, it);
    return LQCD_OK;
}
int blas_nr_update(lqcd_ctx *ctx, cplx *res, const cplx *q, cplx *x, const cplx *p, size_t n) {
    LAUNCH(ctx, k_nr_update, n, res, q, x, p, n, ctx->red);
    return LQCD_OK;
}
int blas_xpby_state(lqcd_ctx *ctx, cplx *p, const cplx *q, size_t n) {
    LAUNCH(ctx, k_xpby_state, n, p, q, n, ctx->red.st);
    return LQCD_OK;
}
int blas_bi_s(lqcd_ctx *ctx, cplx *s, const cplx *r, const cplx *v, size_t n) {
    LAUNCH(ctx, k_bi_s, n, s, r, v, n, ctx->red.st);
    return LQCD_OK;
}
int blas_bi_xr(lqcd_ctx *ctx, cplx *x, const cplx *p, const cplx *s, cplx *r, const cplx *t, const cplx *r0, size_t n) {
    LAUNCH(ctx, k_bi_xr, n, x, p, s, r, t, r0, n, ctx->red);
    return LQCD_OK;
}
int blas_bi_p(lqcd_ctx *ctx, cplx *p, const cplx *r, const cplx *v, size_t n) {
    LAUNCH(ctx, k_bi_p, n, p, r, v, n, ctx->red.st);
    return LQCD_OK;
}
int blas_ms_update_xp(lqcd_ctx *ctx, const MSPtrs &P, const cplx *r, size_t n, int it) {
    LAUNCH(ctx, k_ms_update_xp, n, P, r, n, ctx->red.st, it);
    return LQCD_OK;
}

// ---- exported BLAS ----------------------------------------------------------------------------------
static int check2(const lqcd_ctx *ctx, const lqcd_fermion *a, const lqcd_fermion *b) {
    if (!ctx || !a || !b) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (a->owner != ctx || b->owner != ctx) return lqcd_fail(ctx, LQCD_ERR_ARG, "field belongs to another context");
    if (a->kind != b->kind) return lqcd_fail(ctx, LQCD_ERR_ARG, "fermion kind mismatch");
    return LQCD_OK;
}
static inline size_t flen(const lqcd_ctx *ctx, const lqcd_fermion *f) { return (size_t)ctx->g.nblk * f->ncomp * 32; }

extern "C" int lqcd_blas_axpy(lqcd_ctx *ctx, double are, double aim, const lqcd_fermion *x, lqcd_fermion *y) {
    LQCD_TRY(check2(ctx, x, y));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    size_t n = flen(ctx, x);
    LAUNCH(ctx, k_axpy, n, make_double2(are, aim), x->d, y->d, n);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}
extern "C" int lqcd_blas_xpby(lqcd_ctx *ctx, const lqcd_fermion *x, double bre, double bim, lqcd_fermion *y) {
    LQCD_TRY(check2(ctx, x, y));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    size_t n = flen(ctx, x);
    LAUNCH(ctx, k_xpby, n, x->d, make_double2(bre, bim), y->d, n);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}
extern "C" int lqcd_blas_scale(lqcd_ctx *ctx, double are, double aim, lqcd_fermion *x) {
    LQCD_TRY(check2(ctx, x, x));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    size_t n = flen(ctx, x);
    LAUNCH(ctx, k_scale, n, make_double2(are, aim), x->d, n);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

int comm_allreduce_sum(lqcd_ctx *ctx, double *vals, int n);   // comm.cu (host-side result reduction)

extern "C" int lqcd_blas_dot(lqcd_ctx *ctx, const lqcd_fermion *a, const lqcd_fermion *b, double out[2]) {
    LQCD_TRY(check2(ctx, a, b));
    if (!out) return lqcd_fail(ctx, LQCD_ERR_ARG, "null out");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LQCD_TRY(blas_dot_async(ctx, a->d, b->d, flen(ctx, a), FIN_STORE));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st_host, ctx->red.st, sizeof(SolverState), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double v[2] = {ctx->st_host->red[0], ctx->st_host->red[1]};
    LQCD_TRY(comm_allreduce_sum(ctx, v, 2));
    out[0] = v[0]; out[1] = v[1];
    return LQCD_OK;
}
extern "C" int lqcd_blas_norm2(lqcd_ctx *ctx, const lqcd_fermion *a, double *out) {
    LQCD_TRY(check2(ctx, a, a));
    if (!out) return lqcd_fail(ctx, LQCD_ERR_ARG, "null out");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LQCD_TRY(blas_dot_async(ctx, a->d, a->d, flen(ctx, a), FIN_STORE));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st_host, ctx->red.st, sizeof(SolverState), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double v[1] = {ctx->st_host->red[2]};
    LQCD_TRY(comm_allreduce_sum(ctx, v, 1));
    *out = v[0];
    return LQCD_OK;
}
