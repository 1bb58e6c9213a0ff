// solvers.cu -- operator application entry point and the Krylov loops.
//
// Replaces LatticeDiracOperators.jl's solve_DinvX!(y, A, x) and the cg / "bicg" (CGNR) / bicgstab /
// shiftedcg routines behind it.  Reference call sites: inside calc_UdSfdU!
// (src/md/AbstractMD.jl:129) and evaluate_FermiAction (src/updates/standardHMC.jl:69-71) with A = DdagD,
// and the measurement solves (measure_Pion_correlator.jl:399, measure_chiral_condensate.jl:182) with A = D.
//
// Stopping rule is the reference's: real(r.r) < eps_CG, ABSOLUTE and SQUARED (default 1e-19,
// src/system/parameter_structs.jl:174); the solution vector doubles as the initial guess; running out of
// MaxCGstep is an error (LQCD_ERR_NOCONV).
//
// Device-resident loop: every scalar of the recurrence lives in SolverState (device memory) and is
// produced by the "finish" step of the reducing kernel that computes the dot product
// (reduce.cuh).  The host only enqueues kernels; it polls a pinned copy of the state one batch
// behind the GPU, so the GPU never waits for the host.  Kernels launched past convergence are no-ops.
#include "lqcd_internal.cuh"

int blas_dot_async(lqcd_ctx *ctx, const cplx *a, const cplx *b, size_t n, int finish);
int blas_resid_init(lqcd_ctx *ctx, const cplx *b, const cplx *q, cplx *r, cplx *p, cplx *r0, size_t n, int finish);
int blas_cg_update_r(lqcd_ctx *ctx, cplx *r, const cplx *q, size_t n, int finish);
int blas_cg_update_xp(lqcd_ctx *ctx, cplx *x, cplx *p, const cplx *r, size_t n, int it);
int blas_nr_update(lqcd_ctx *ctx, cplx *res, const cplx *q, cplx *x, const cplx *p, size_t n);
int blas_xpby_state(lqcd_ctx *ctx, cplx *p, const cplx *q, size_t n);
int blas_bi_s(lqcd_ctx *ctx, cplx *s, const cplx *r, const cplx *v, size_t n);
int blas_bi_xr(lqcd_ctx *ctx, cplx *x, const cplx *p, const cplx *s, cplx *r, const cplx *t, const cplx *r0, size_t n);
int blas_bi_p(lqcd_ctx *ctx, cplx *p, const cplx *r, const cplx *v, size_t n);
int blas_ms_update_xp(lqcd_ctx *ctx, const MSPtrs &P, const cplx *r, size_t n, int it);
int comm_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse);   // comm.cu
int comm_check_error(lqcd_ctx *ctx);
int comm_timing_report(lqcd_ctx *ctx, const char *what);

static int check_op(const lqcd_ctx *ctx, const lqcd_op *op) {
    if (!op) return lqcd_fail(ctx, LQCD_ERR_ARG, "null operator descriptor");
    if (op->kind != LQCD_WILSON && op->kind != LQCD_STAGGERED) return lqcd_fail(ctx, LQCD_ERR_ARG, "unknown operator kind %d", op->kind);
    for (int i = 0; i < 4; i++)
        if (op->bc[i] != 1.0 && op->bc[i] != -1.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "boundary phase bc[%d] = %g must be +-1", i, op->bc[i]);
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "operator applied before lqcd_gauge_upload");
    return LQCD_OK;
}

int eo_mhat(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse);   // wilson_eo.cu
int wilson_dslash_general_r(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse);   // wilson_general_r.cu

static int one_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse) {
    if (ctx->eo_active) return eo_mhat(ctx, op, y, x, dagger, fuse);      // lqcd_solve_eo: the "operator" is Mhat on even half fields
    if (op->kind == LQCD_WILSON && op->r != 1.0) return wilson_dslash_general_r(ctx, op, y, x, dagger, fuse);      // rare general case
    if (ctx->nranks > 1) return comm_dslash(ctx, op, y, x, dagger, fuse);
    if (op->kind == LQCD_WILSON) return launch_wilson_dslash(ctx, op, y, x, dagger, fuse, ctx->stream);
    return launch_staggered_dslash(ctx, op, y, x, dagger, fuse, ctx->stream);
}

// y = D x | D^dag x | D^dag D x (through tmp).  `first` / `last` are the fused epilogues of the first /
// last kernel (for D / D^dag only `last` is used).
static int apply_async(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int mode, cplx *tmp,
                       const DslashFuse *first, const DslashFuse *last) {
    if (mode == LQCD_OP_D) return one_dslash(ctx, op, y, x, 0, last);
    if (mode == LQCD_OP_DDAG) return one_dslash(ctx, op, y, x, 1, last);
    LQCD_TRY(one_dslash(ctx, op, tmp, x, 0, first));
    return one_dslash(ctx, op, y, tmp, 1, last);
}

extern "C" int lqcd_dslash(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *x, int mode) {
    if (!ctx || !y || !x) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    LQCD_TRY(check_op(ctx, op));
    if (y->owner != ctx || x->owner != ctx) return lqcd_fail(ctx, LQCD_ERR_ARG, "field belongs to another context");
    if (y->kind != op->kind || x->kind != op->kind) return lqcd_fail(ctx, LQCD_ERR_ARG, "fermion kind does not match the operator");
    if (mode < 0 || mode > 2) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad mode %d", mode);
    if (y == x) return lqcd_fail(ctx, LQCD_ERR_ARG, "mul!: output aliases input");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    lqcd_fermion *tmp = nullptr;
    if (mode == LQCD_OP_DDAGD) LQCD_TRY(get_scratch(ctx, op->kind, 0, &tmp));
    LQCD_TRY(apply_async(ctx, op, y->d, x->d, mode, tmp ? tmp->d : nullptr, nullptr, nullptr));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return comm_check_error(ctx);
}

// ---- solver driver -----------------------------------------------------------------------------------
static int state_init(lqcd_ctx *ctx, double eps, int maxsteps, const double *shifts, int nshift, double **hist_dev) {
    SolverState *h = &ctx->st_host[0];
    memset(h, 0, sizeof *h);
    h->eps = eps; h->maxit = maxsteps; h->nshift = nshift;
    h->alpha_old = 1.0; h->beta_old = 0.0;
    for (int j = 0; j < nshift; j++) { h->shift[j] = shifts[j]; h->zeta[j] = h->zeta_old[j] = 1.0; }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->red.st, h, sizeof *h, cudaMemcpyHostToDevice, ctx->stream));
    if (hist_dev) {
        if (ctx->hist_cap < maxsteps + 1) {
            if (ctx->hist_dev) CUDA_TRY(ctx, cudaFree(ctx->hist_dev));
            ctx->hist_dev = nullptr; ctx->hist_cap = 0;
            CUDA_TRY(ctx, cudaMalloc(&ctx->hist_dev, sizeof(double) * (maxsteps + 1)));
            ctx->hist_cap = maxsteps + 1;
        }
        *hist_dev = ctx->hist_dev;
    }
    return LQCD_OK;
}

// Polls the device state one batch behind the enqueue front.  body(it) enqueues iteration `it`.
template <class Body>
static int run_loop(lqcd_ctx *ctx, int maxsteps, Body body, int *iters, double *resid_sq) {
    int batch = 8;
    if (const char *e = getenv("LQCD_CG_BATCH")) { int v = atoi(e); if (v >= 1 && v <= 1024) batch = v; }
    cudaEvent_t ev[2] = {ctx->ev_poll[0], ctx->ev_poll[1]};
    int nb = 0;                 // batches enqueued
    bool done = false;
    SolverState fin;
    // state after the init kernels
    CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->st_host[1], ctx->red.st, sizeof(SolverState), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ev[0], ctx->stream));
    int it = 0;
    int pending = 1;            // slot (1 or 2) holding the newest enqueued snapshot
    while (true) {
        // enqueue next batch (unless everything is already enqueued)
        int prev_slot = pending;
        bool enq = false;
        if (it < maxsteps) {
            int hi = it + batch < maxsteps ? it + batch : maxsteps;
            for (int i = it + 1; i <= hi; i++) LQCD_TRY(body(i));
            it = hi; nb++; enq = true;
            pending = (pending == 1) ? 2 : 1;
            CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->st_host[pending], ctx->red.st, sizeof(SolverState), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaEventRecord(ev[pending - 1], ctx->stream));
        }
        // wait for the snapshot taken BEFORE this batch
        CUDA_TRY(ctx, cudaEventSynchronize(ev[prev_slot - 1]));
        fin = ctx->st_host[prev_slot];
        if (fin.done) { done = true; break; }
        if (!enq) break;        // nothing more to enqueue and the last snapshot is not converged
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (!done) {                // the very last snapshot may hold the converged state
        fin = ctx->st_host[pending];
        done = fin.done != 0;
    }
    if (iters) *iters = done ? fin.iters : fin.it;
    if (resid_sq) *resid_sq = fin.rr;
    (void)nb;
    LQCD_TRY(comm_check_error(ctx));
    if (done && fin.failed == 2) return lqcd_fail(ctx, LQCD_ERR_COMM, "all-reduce timed out on the device at step %d", fin.it);
    if (done && fin.failed)
        return lqcd_fail(ctx, LQCD_ERR_NOCONV, "Krylov breakdown: |r|^2 is not finite at step %d (e.g. BiCGStab with r0~ = r0 on a point source)", fin.it);
    if (!done)
        return lqcd_fail(ctx, LQCD_ERR_NOCONV, "solver not converged after %d steps (|r|^2 = %.6e, eps = %.3e)", fin.it, fin.rr, fin.eps);
    return LQCD_OK;
}

static inline size_t flen(const lqcd_ctx *ctx, const lqcd_fermion *f) { return (size_t)ctx->g.nblk * f->ncomp * 32; }

int solve_impl(lqcd_ctx *ctx, const lqcd_op *op, cplx *x, const cplx *bb, size_t n, int method, int target,
               double eps, int maxsteps, int *iters, double *resid_sq, double *hist);

extern "C" int lqcd_solve(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *b, int method, int target,
                          double eps, int maxsteps, int *iters, double *resid_sq, double *hist) {
    if (!ctx || !y || !b) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    LQCD_TRY(check_op(ctx, op));
    if (y->owner != ctx || b->owner != ctx) return lqcd_fail(ctx, LQCD_ERR_ARG, "field belongs to another context");
    if (y->kind != op->kind || b->kind != op->kind) return lqcd_fail(ctx, LQCD_ERR_ARG, "fermion kind does not match the operator");
    if (y == b) return lqcd_fail(ctx, LQCD_ERR_ARG, "solve_DinvX!: solution aliases the source");
    if (maxsteps < 1 || !(eps >= 0.0)) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad eps / maxsteps");
    if (method == LQCD_SOLVER_CG && target != LQCD_OP_DDAGD) return lqcd_fail(ctx, LQCD_ERR_ARG, "CG needs the Hermitian target DdagD");
    if (method != LQCD_SOLVER_CG && target != LQCD_OP_D && target != LQCD_OP_DDAG) return lqcd_fail(ctx, LQCD_ERR_ARG, "CGNR/BiCGStab solve D or D^dag");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->stag_even_solve && op->kind == LQCD_STAGGERED && method == LQCD_SOLVER_CG && !hist && ctx->nranks == 1)
        return lqcd_solve_staggered_even(ctx, op, y, b, eps, maxsteps, iters, resid_sq);      // staggered_eo.cu (off unless switched on)
    return solve_impl(ctx, op, y->d, b->d, flen(ctx, y), method, target, eps, maxsteps, iters, resid_sq, hist);
}

// The Krylov loops on raw device vectors of n complex numbers (full fields from lqcd_solve, even half fields from
// lqcd_solve_eo, where apply_async applies Mhat).
int solve_impl(lqcd_ctx *ctx, const lqcd_op *op, cplx *x, const cplx *bb, size_t n, int method, int target,
               double eps, int maxsteps, int *iters, double *resid_sq, double *hist) {
    const int kind = op->kind;
    double *hd = nullptr;
    LQCD_TRY(state_init(ctx, eps, maxsteps, nullptr, 0, hist ? &hd : nullptr));
    ctx->red.hist = hd;
    DslashFuse plain = DslashFuse(); plain.use_state = 1;
    int rc = LQCD_OK;

    if (method == LQCD_SOLVER_CG) {
        lqcd_fermion *fr, *fp, *fq, *ft;
        LQCD_TRY(get_scratch(ctx, kind, 0, &ft)); LQCD_TRY(get_scratch(ctx, kind, 1, &fr));
        LQCD_TRY(get_scratch(ctx, kind, 2, &fp)); LQCD_TRY(get_scratch(ctx, kind, 3, &fq));
        cplx *r = fr->d, *p = fp->d, *q = fq->d, *t = ft->d;
        LQCD_TRY(apply_async(ctx, op, q, x, LQCD_OP_DDAGD, t, nullptr, nullptr));
        LQCD_TRY(blas_resid_init(ctx, bb, q, r, p, nullptr, n, FIN_CG_INIT));
        // Per iteration: 3 kernels, 3072 B/site.  <p, D^dag D p> = |D p|^2 is reduced in the epilogue of the FIRST
        // Dslash (no extra read of p), so alpha is known before the second Dslash starts; that kernel then applies
        // r <- r - alpha (D^dag t) in its epilogue and reduces |r|^2 -- q = D^dag D p is never written to memory.
        // LQCD_CG_FUSE=0 selects the 4-kernel form (explicit <p,q> dot fused in the second Dslash, separate r update).
        static int cgfuse = -1;
        if (cgfuse < 0) { const char *e = getenv("LQCD_CG_FUSE"); cgfuse = (e && atoi(e) == 0) ? 0 : 1; }
        DslashFuse pq = DslashFuse(); pq.use_state = 1; pq.dot_with = p; pq.finish = FIN_CG_PQ;
        DslashFuse f1 = DslashFuse(); f1.use_state = 1; f1.want_norm = 1; f1.finish = FIN_CG_PQN;
        DslashFuse f2 = DslashFuse(); f2.use_state = 1; f2.want_norm = 1; f2.finish = FIN_CG_RRN; f2.axpy_r = r;
        rc = run_loop(ctx, maxsteps, [&](int it) -> int {
            if (cgfuse) {
                LQCD_TRY(apply_async(ctx, op, q, p, LQCD_OP_DDAGD, t, &f1, &f2));     // t = D p (alpha); r -= alpha D^dag t (|r|^2, beta)
            } else {
                LQCD_TRY(apply_async(ctx, op, q, p, LQCD_OP_DDAGD, t, &plain, &pq));  // q = D^dag D p, pq = <p,q>
                LQCD_TRY(blas_cg_update_r(ctx, r, q, n, FIN_CG_RR));                  // r -= alpha q, |r|^2, beta
            }
            return blas_cg_update_xp(ctx, x, p, r, n, it);                            // x += alpha p, p = r + beta p
        }, iters, resid_sq);
    } else if (method == LQCD_SOLVER_CGNR) {
        const int A = target, Ad = (target == LQCD_OP_D) ? LQCD_OP_DDAG : LQCD_OP_D;
        lqcd_fermion *fres, *fp, *fq;
        LQCD_TRY(get_scratch(ctx, kind, 0, &fres)); LQCD_TRY(get_scratch(ctx, kind, 1, &fp)); LQCD_TRY(get_scratch(ctx, kind, 2, &fq));
        cplx *res = fres->d, *p = fp->d, *q = fq->d;
        LQCD_TRY(apply_async(ctx, op, q, x, A, nullptr, nullptr, nullptr));
        LQCD_TRY(blas_resid_init(ctx, bb, q, res, nullptr, nullptr, n, FIN_CG_INIT));
        DslashFuse c1 = DslashFuse(); c1.use_state = 1; c1.want_norm = 1; c1.finish = FIN_NR_C1;
        DslashFuse c2 = c1; c2.finish = FIN_NR_C2;
        DslashFuse c3 = c1; c3.finish = FIN_NR_C3;
        LQCD_TRY(apply_async(ctx, op, q, res, Ad, nullptr, nullptr, &c1));            // q = A^dag res, c1 = |q|^2
        LQCD_TRY(blas_copy(ctx, p, q, n));
        rc = run_loop(ctx, maxsteps, [&](int) -> int {
            LQCD_TRY(apply_async(ctx, op, q, p, A, nullptr, nullptr, &c2));           // q = A p, alpha = c1/|q|^2
            LQCD_TRY(blas_nr_update(ctx, res, q, x, p, n));                           // res -= alpha q, x += alpha p, |res|^2
            LQCD_TRY(apply_async(ctx, op, q, res, Ad, nullptr, nullptr, &c3));        // q = A^dag res, beta = |q|^2/c1
            return blas_xpby_state(ctx, p, q, n);                                     // p = beta p + q
        }, iters, resid_sq);
    } else if (method == LQCD_SOLVER_BICGSTAB) {
        const int A = target;
        lqcd_fermion *fr, *fr0, *fp, *fv, *fs, *ft;
        LQCD_TRY(get_scratch(ctx, kind, 0, &fr)); LQCD_TRY(get_scratch(ctx, kind, 1, &fr0)); LQCD_TRY(get_scratch(ctx, kind, 2, &fp));
        LQCD_TRY(get_scratch(ctx, kind, 3, &fv)); LQCD_TRY(get_scratch(ctx, kind, 4, &fs)); LQCD_TRY(get_scratch(ctx, kind, 5, &ft));
        cplx *r = fr->d, *r0 = fr0->d, *p = fp->d, *v = fv->d, *s = fs->d, *t = ft->d;
        LQCD_TRY(apply_async(ctx, op, v, x, A, nullptr, nullptr, nullptr));
        LQCD_TRY(blas_resid_init(ctx, bb, v, r, p, r0, n, FIN_BI_INIT));
        DslashFuse fa = DslashFuse(); fa.use_state = 1; fa.dot_with = r0; fa.finish = FIN_BI_ALPHA;
        DslashFuse fo = DslashFuse(); fo.use_state = 1; fo.dot_with = s; fo.finish = FIN_BI_OMEGA;
        rc = run_loop(ctx, maxsteps, [&](int) -> int {
            LQCD_TRY(apply_async(ctx, op, v, p, A, nullptr, nullptr, &fa));           // v = A p, alpha = rho/<r0,v>
            LQCD_TRY(blas_bi_s(ctx, s, r, v, n));                                     // s = r - alpha v
            LQCD_TRY(apply_async(ctx, op, t, s, A, nullptr, nullptr, &fo));           // t = A s, omega = <t,s>/|t|^2
            LQCD_TRY(blas_bi_xr(ctx, x, p, s, r, t, r0, n));                          // x, r, |r|^2, rho', beta
            return blas_bi_p(ctx, p, r, v, n);                                        // p = r + beta (p - omega v)
        }, iters, resid_sq);
    } else {
        ctx->red.hist = nullptr;
        return lqcd_fail(ctx, LQCD_ERR_ARG, "unknown solver method %d", method);
    }
    ctx->red.hist = nullptr;
    if (hist && hd) {
        int nh = (iters ? *iters : maxsteps) + 1;
        if (nh > maxsteps + 1) nh = maxsteps + 1;
        CUDA_TRY(ctx, cudaMemcpy(hist, hd, sizeof(double) * nh, cudaMemcpyDeviceToHost));
    }
    return rc;
}

extern "C" int lqcd_multishift_cg(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *const ys[], const lqcd_fermion *b,
                                  const double *shifts, int nshift, double eps, int maxsteps, int *iters, double *resid_sq) {
    if (!ctx || !ys || !b || !shifts) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    LQCD_TRY(check_op(ctx, op));
    if (nshift < 1 || nshift > LQCD_MAX_SHIFTS) return lqcd_fail(ctx, LQCD_ERR_ARG, "nshift must be in [1, %d]", LQCD_MAX_SHIFTS);
    for (int j = 1; j < nshift; j++)
        if (shifts[j] < shifts[0]) return lqcd_fail(ctx, LQCD_ERR_ARG, "shifts[0] must be the smallest shift");
    if (b->owner != ctx || b->kind != op->kind) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad source field");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int kind = op->kind;
    const size_t n = flen(ctx, b);
    MSPtrs P;
    memset(&P, 0, sizeof P);
    lqcd_fermion *fr, *fq, *ft;
    LQCD_TRY(get_scratch(ctx, kind, 0, &ft)); LQCD_TRY(get_scratch(ctx, kind, 1, &fr)); LQCD_TRY(get_scratch(ctx, kind, 2, &fq));
    for (int j = 0; j < nshift; j++) {
        if (!ys[j] || ys[j]->owner != ctx || ys[j]->kind != kind || ys[j] == b) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad solution field %d", j);
        lqcd_fermion *fp;
        LQCD_TRY(get_scratch(ctx, kind, 3 + j, &fp));
        P.x[j] = ys[j]->d; P.p[j] = fp->d;
    }
    LQCD_TRY(state_init(ctx, eps, maxsteps, shifts, nshift, nullptr));
    ctx->red.hist = nullptr;
    cplx *r = fr->d, *q = fq->d, *t = ft->d;
    for (int j = 0; j < nshift; j++) {
        LQCD_TRY(blas_zero(ctx, P.x[j], n));
        LQCD_TRY(blas_copy(ctx, P.p[j], b->d, n));
    }
    LQCD_TRY(blas_copy(ctx, r, b->d, n));
    LQCD_TRY(blas_dot_async(ctx, b->d, b->d, n, FIN_CG_INIT));
    DslashFuse plain = DslashFuse(); plain.use_state = 1;
    DslashFuse pq = DslashFuse(); pq.use_state = 1; pq.dot_with = P.p[0]; pq.finish = FIN_MS_PQ;
    pq.shift = shifts[0]; pq.shift_src = P.p[0];
    return run_loop(ctx, maxsteps, [&](int it) -> int {
        LQCD_TRY(apply_async(ctx, op, q, P.p[0], LQCD_OP_DDAGD, t, &plain, &pq));     // q = (D^dag D + s0) p0, alphas, zetas
        LQCD_TRY(blas_cg_update_r(ctx, r, q, n, FIN_MS_RR));                          // r -= alpha q, |r|^2, betas
        return blas_ms_update_xp(ctx, P, r, n, it);                                   // x_j, p_j for all shifts
    }, iters, resid_sq);
}

// ---- timing helper used by bench.py (CUDA events on the library's own stream) ----------------------
extern "C" int lqcd_time_dslash(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *x, int mode,
                                int reps, int flush_l2, double *ms_mean, double *ms_min) {
    if (!ctx || !y || !x || reps < 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad argument");
    LQCD_TRY(check_op(ctx, op));
    if (y->kind != op->kind || x->kind != op->kind || y == x) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad fields");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    lqcd_fermion *tmp = nullptr;
    if (mode == LQCD_OP_DDAGD) LQCD_TRY(get_scratch(ctx, op->kind, 0, &tmp));
    if (flush_l2 && !ctx->flush) {
        ctx->flush_bytes = (size_t)512 << 20;
        CUDA_TRY(ctx, cudaMalloc(&ctx->flush, ctx->flush_bytes));
    }
    double sum = 0.0, mn = 1e300;
    if (!flush_l2) {     // back-to-back: ONE event bracket around all applications (no host sync inside -> no rank skew)
        // two untimed applications first: across ranks they align the GPUs through the halo flags (a face tile of application k
        // waits for the neighbours' application k), so the bracket below does not measure the launch skew left by a host barrier
        for (int i = 0; i < 2; i++)
            LQCD_TRY(apply_async(ctx, op, y->d, x->d, mode, tmp ? tmp->d : nullptr, nullptr, nullptr));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        for (int i = 0; i < reps; i++)
            LQCD_TRY(apply_async(ctx, op, y->d, x->d, mode, tmp ? tmp->d : nullptr, nullptr, nullptr));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (ms_mean) *ms_mean = ms / reps;
        if (ms_min) *ms_min = ms / reps;
        if (ctx->nranks > 1) LQCD_TRY(comm_timing_report(ctx, op->kind == LQCD_WILSON ? "Wilson Dslash" : "staggered Dslash"));
        return comm_check_error(ctx);
    }
    for (int i = 0; i < reps; i++) {
        if (flush_l2) CUDA_TRY(ctx, cudaMemsetAsync(ctx->flush, i & 0xff, ctx->flush_bytes, ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        LQCD_TRY(apply_async(ctx, op, y->d, x->d, mode, tmp ? tmp->d : nullptr, nullptr, nullptr));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        sum += ms; if (ms < mn) mn = ms;
    }
    if (ms_mean) *ms_mean = sum / reps;
    if (ms_min) *ms_min = mn;
    return comm_check_error(ctx);
}
