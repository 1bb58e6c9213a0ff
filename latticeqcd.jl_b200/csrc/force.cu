// force.cu -- pseudofermion MD force  U_mu dS_f/dU_mu  on the device.
//
// Replaces LatticeDiracOperators.jl's calc_UdSfdU!(UdSfdU, fermi_action, U, eta) called from
// P_update_fermion! (src/md/AbstractMD.jl:120-135; stout variant src/md/standardMD.jl:192-227):
//     X = (D^dag D)^-1 eta   (CG, solvers.cu)        Y = D X
//     Wilson   : F_mu(n) = -kappa sum_s [U_mu(n) P- X(n+mu)]_s (x) conj([P- Y(n)]_s)
//                          +kappa sum_s [P+ X(n)]_s (x) conj([U_mu(n) P+ Y(n+mu)]_s)          (P+- = 1 +- gamma_mu)
//     staggered: F_mu(n) = (1/2) eta_mu(n) { [U X(n+mu)] (x) conj(Y(n)) + X(n) (x) conj([U Y(n+mu)]) }
// such that dS_f/d eps = -2 Re tr[A F_mu(n)] for U_mu(n) -> exp(eps A) U_mu(n) with S_f = eta^dag (D^dag D)^-1 eta
// (the identity is the finite-difference test in tests/).  The spin trace uses the
// rank-2 structure of the projectors: tr_spin[(1-g) v w^dag] = sum_{s=0,1} (P v)_s conj((P w)_s).
// The reference then applies p_mu += -eps dtau * Traceless_antihermitian(F_mu) on the CPU (AbstractMD.jl:131-133),
// so the result is returned in the host link layout.
//
// Multi-GPU (one process per GPU): the CG and Y = D X run through comm.cu as usual; the outer products need X(n+mu) and
// Y(n+mu) of the upper neighbour rank for the sites of the HIGH faces.  force_pack_kernel stores the spin-projected halves
// of this rank's LOW faces (Wilson: P- X | P+ Y = 12 complex per face site; staggered: X | Y) straight into the lower
// neighbours' force slots over NVLink and raises one sequence flag per direction; the MULTI force kernels wait for their own
// flags (clock64 timeout -> LQCD_ERR_COMM) and read the slots with ld.global.cg.  Slot reuse is ordered by the global
// <eta, X> reduction that closes every call: its in-kernel all-reduce completes on a rank only after EVERY rank has posted
// its contribution, which each rank does after its own force kernel (stream order), so the next call's pack can never
// overwrite a slot that a neighbour is still reading.
#include "lqcd_internal.cuh"
#include "wilson_spin.cuh"
#include "site_map.cuh"
#include "halo_pack.cuh"
#include <cstring>

struct ForceArgs {
    cplx *out;               // device link layout [((blk*4+mu)*9 + a*3+b)*32 + lane]
    const cplx *X, *Y, *gauge;
    Geom g;
    double kappa;
    double bc[4];
    double coef;             // out = [out +] coef * force
    int accumulate;
    ForceHalo fh;            // MULTI kernels only
};

// all threads of a MULTI force CTA: wait until the upper neighbours' low faces of this call have landed
__device__ __forceinline__ void wait_force_flags(const ForceArgs &A) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        bool good = true;
        for (int m = 0; m < 4 && good; m++) {
            if (!A.g.part[m]) continue;
            while (ld_acquire_sys(A.fh.recv_flag[m]) < A.fh.seq)
                if (clock64() - t0 > A.fh.timeout_cycles) { good = false; *A.fh.err = 3000000 + m * 100000 + (int)(A.fh.seq % 100000); break; }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void load_spinor(cplx (&p)[12], const cplx *f, int s) {
    const cplx *sp = f + (size_t)(s >> 5) * (12 * 32) + (s & 31);
#pragma unroll
    for (int k = 0; k < 12; k++) p[k] = ldg128(sp + k * 32);
}

template <int MU, int MULTI>
__device__ __forceinline__ void wilson_force_dir(const ForceArgs &A, int s, int coord, int dim, int stride,
                                                 const cplx (&Xn)[12], const cplx (&Yn)[12], int x, int y, int z, int t) {
    const bool w = (coord == dim - 1);
    const int ns = w ? s - (dim - 1) * stride : s + stride;
    cplx hx0[3], hx1[3], hy0[3], hy1[3], px0[3], px1[3], py0[3], py1[3];
    double phase;
    if (MULTI && w && A.g.part[MU]) {          // neighbour is on the upper rank: projected halves from the force slot
        const int f = face_index<MU>(A.g, x, y, z, t);
        const cplx *src = A.fh.recv[MU] + (size_t)(f >> 5) * (12 * 32) + (f & 31);
        phase = A.fh.plast[MU] ? A.bc[MU] : 1.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            hx0[c] = __ldcg(src + c * 32);       hx1[c] = __ldcg(src + (3 + c) * 32);
            hy0[c] = __ldcg(src + (6 + c) * 32); hy1[c] = __ldcg(src + (9 + c) * 32);
        }
    } else {
        phase = w ? A.bc[MU] : 1.0;
        cplx Xf[12], Yf[12];
        load_spinor(Xf, A.X, ns);
        load_spinor(Yf, A.Y, ns);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            project<MU, -1>(hx0[c], hx1[c], Xf[c], Xf[3 + c], Xf[6 + c], Xf[9 + c]);      // P- X(n+mu)
            project<MU, +1>(hy0[c], hy1[c], Yf[c], Yf[3 + c], Yf[6 + c], Yf[9 + c]);      // P+ Y(n+mu)
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        project<MU, -1>(py0[c], py1[c], Yn[c], Yn[3 + c], Yn[6 + c], Yn[9 + c]);      // P- Y(n)
        project<MU, +1>(px0[c], px1[c], Xn[c], Xn[3 + c], Xn[6 + c], Xn[9 + c]);      // P+ X(n)
        hx0[c] = cscale(phase, hx0[c]); hx1[c] = cscale(phase, hx1[c]);
        hy0[c] = cscale(phase, hy0[c]); hy1[c] = cscale(phase, hy1[c]);
    }
    const cplx *lk = A.gauge + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
    cplx gx0[3], gx1[3], gy0[3], gy1[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        gx0[a] = gx1[a] = gy0[a] = gy1[a] = cmake(0, 0);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            cplx u = ldg128(lk + (a * 3 + b) * 32);
            cfma(gx0[a], u, hx0[b]); cfma(gx1[a], u, hx1[b]);
            cfma(gy0[a], u, hy0[b]); cfma(gy1[a], u, hy1[b]);
        }
    }
    cplx *o = A.out + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            // t1 = sum_s gx_s[a] conj(py_s[b]) ; t2 = sum_s px_s[a] conj(gy_s[b])      (x conj(y) = conj(y) x)
            cplx t1 = cmake(0, 0), t2 = cmake(0, 0);
            cfmac(t1, py0[b], gx0[a]); cfmac(t1, py1[b], gx1[a]);
            cfmac(t2, gy0[b], px0[a]); cfmac(t2, gy1[b], px1[a]);
            const double ck = A.coef * A.kappa;
            cplx v = cmake(ck * (t2.x - t1.x), ck * (t2.y - t1.y));
            if (A.accumulate) v = cadd(v, o[(a * 3 + b) * 32]);
            o[(a * 3 + b) * 32] = v;
        }
}

template <int MULTI>
__global__ void __launch_bounds__(128) wilson_force_kernel(const ForceArgs A) {
    if (MULTI) wait_force_flags(A);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.g.V) return;
    int x, y, z, t;
    site_coords(A.g, s, x, y, z, t);
    cplx Xn[12], Yn[12];
    load_spinor(Xn, A.X, s);
    load_spinor(Yn, A.Y, s);
    wilson_force_dir<0, MULTI>(A, s, x, A.g.X, 1, Xn, Yn, x, y, z, t);
    wilson_force_dir<1, MULTI>(A, s, y, A.g.Y, A.g.X, Xn, Yn, x, y, z, t);
    wilson_force_dir<2, MULTI>(A, s, z, A.g.Z, A.g.X * A.g.Y, Xn, Yn, x, y, z, t);
    wilson_force_dir<3, MULTI>(A, s, t, A.g.T, A.g.X * A.g.Y * A.g.Z, Xn, Yn, x, y, z, t);
}

template <int MU, int MULTI>
__device__ __forceinline__ void stag_force_dir(const ForceArgs &A, int s, int coord, int dim, int stride, double eta,
                                               const cplx (&Xn)[3], const cplx (&Yn)[3], int x, int y, int z, int t) {
    const bool w = (coord == dim - 1);
    const int ns = w ? s - (dim - 1) * stride : s + stride;
    cplx xf[3], yf[3], hx[3], hy[3];
    if (MULTI && w && A.g.part[MU]) {
        const int f = face_index<MU>(A.g, x, y, z, t);
        const cplx *src = A.fh.recv[MU] + (size_t)(f >> 5) * (12 * 32) + (f & 31);
        const double phase = A.fh.plast[MU] ? A.bc[MU] : 1.0;
#pragma unroll
        for (int c = 0; c < 3; c++) { xf[c] = cscale(phase, __ldcg(src + c * 32)); yf[c] = cscale(phase, __ldcg(src + (3 + c) * 32)); }
    } else {
        const double phase = w ? A.bc[MU] : 1.0;
        const cplx *xp = A.X + (size_t)(ns >> 5) * (3 * 32) + (ns & 31), *yp = A.Y + (size_t)(ns >> 5) * (3 * 32) + (ns & 31);
#pragma unroll
        for (int c = 0; c < 3; c++) { xf[c] = cscale(phase, ldg128(xp + c * 32)); yf[c] = cscale(phase, ldg128(yp + c * 32)); }
    }
    const cplx *lk = A.gauge + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
#pragma unroll
    for (int a = 0; a < 3; a++) {
        hx[a] = hy[a] = cmake(0, 0);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            cplx u = ldg128(lk + (a * 3 + b) * 32);
            cfma(hx[a], u, xf[b]); cfma(hy[a], u, yf[b]);
        }
    }
    cplx *o = A.out + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
    const double cf = 0.5 * eta * A.coef;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            cplx t = cmake(0, 0);
            cfmac(t, Yn[b], hx[a]);
            cfmac(t, hy[b], Xn[a]);
            cplx v = cscale(cf, t);
            if (A.accumulate) v = cadd(v, o[(a * 3 + b) * 32]);
            o[(a * 3 + b) * 32] = v;
        }
}

template <int MULTI>
__global__ void __launch_bounds__(128) staggered_force_kernel(const ForceArgs A) {
    if (MULTI) wait_force_flags(A);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.g.V) return;
    int x, y, z, t;
    site_coords(A.g, s, x, y, z, t);
    cplx Xn[3], Yn[3];
    const cplx *xp = A.X + (size_t)(s >> 5) * (3 * 32) + (s & 31), *yp = A.Y + (size_t)(s >> 5) * (3 * 32) + (s & 31);
#pragma unroll
    for (int c = 0; c < 3; c++) { Xn[c] = ldg128(xp + c * 32); Yn[c] = ldg128(yp + c * 32); }
    const int gx = x + A.g.o[0], gy = y + A.g.o[1], gz = z + A.g.o[2];
    stag_force_dir<0, MULTI>(A, s, x, A.g.X, 1, 1.0, Xn, Yn, x, y, z, t);
    stag_force_dir<1, MULTI>(A, s, y, A.g.Y, A.g.X, (gx & 1) ? -1.0 : 1.0, Xn, Yn, x, y, z, t);
    stag_force_dir<2, MULTI>(A, s, z, A.g.Z, A.g.X * A.g.Y, ((gx + gy) & 1) ? -1.0 : 1.0, Xn, Yn, x, y, z, t);
    stag_force_dir<3, MULTI>(A, s, t, A.g.T, A.g.X * A.g.Y * A.g.Z, ((gx + gy + gz) & 1) ? -1.0 : 1.0, Xn, Yn, x, y, z, t);
}

// ---- multi-GPU: ship the projected LOW-face halves of X and Y to the lower neighbours ----------------------------------------
template <int MU>
__device__ __forceinline__ void wilson_force_pack_site(const ForceArgs &A, cplx *dst, int s) {
    cplx Xs[12], Ys[12];
    load_spinor(Xs, A.X, s);
    load_spinor(Ys, A.Y, s);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        cplx a0, a1, b0, b1;
        project<MU, -1>(a0, a1, Xs[c], Xs[3 + c], Xs[6 + c], Xs[9 + c]);      // P- X(m): the receiver's X(n+mu)
        project<MU, +1>(b0, b1, Ys[c], Ys[3 + c], Ys[6 + c], Ys[9 + c]);      // P+ Y(m)
        dst[c * 32] = a0; dst[(3 + c) * 32] = a1; dst[(6 + c) * 32] = b0; dst[(9 + c) * 32] = b1;
    }
}

__global__ void __launch_bounds__(128) force_pack_kernel(const ForceArgs A, int kind) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.fh.start[4]) {
        int mu = 0;
        while (mu < 3 && i >= A.fh.start[mu + 1]) mu++;
        const int f = i - A.fh.start[mu];
        const int s = face_site(A.g, mu, f, 0);
        cplx *dst = A.fh.send[mu] + (size_t)(f >> 5) * (12 * 32) + (f & 31);
        if (kind == LQCD_WILSON) {
            switch (mu) {
            case 0: wilson_force_pack_site<0>(A, dst, s); break;
            case 1: wilson_force_pack_site<1>(A, dst, s); break;
            case 2: wilson_force_pack_site<2>(A, dst, s); break;
            default: wilson_force_pack_site<3>(A, dst, s); break;
            }
        } else {
            const cplx *xp = A.X + (size_t)(s >> 5) * (3 * 32) + (s & 31), *yp = A.Y + (size_t)(s >> 5) * (3 * 32) + (s & 31);
#pragma unroll
            for (int c = 0; c < 3; c++) { dst[c * 32] = ldg128(xp + c * 32); dst[(3 + c) * 32] = ldg128(yp + c * 32); }
        }
    }
    // publish (same pattern as halo_pack_cta): CTA barrier, system fence + ticket, the last CTA raises the flags
    __syncthreads();
    __shared__ int fpack_last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        fpack_last = (atomicInc(A.fh.ticket, gridDim.x - 1) == gridDim.x - 1);
    }
    __syncthreads();
    if (fpack_last && threadIdx.x < 4) {
        __threadfence_system();
        if (A.g.part[threadIdx.x]) st_release_sys(A.fh.send_flag[threadIdx.x], A.fh.seq);
    }
}

int download_links_from(lqcd_ctx *ctx, const cplx *dev_links, double *const U_mu[4], int ndw);     // context.cu
int comm_check_error(lqcd_ctx *ctx);                                                                // comm.cu
int clover_force_accumulate(lqcd_ctx *ctx, const lqcd_op *op, const cplx *X, const cplx *Y, double weight);   // clover_force.cu

// F <- [F +] coef * force(X, Y) into the device-resident link-shaped buffer (RHMC: sum_j alpha_j force(X_j, Y_j) without
// leaving the GPU).  Collective across ranks.  `barrier_with`: field whose <.,X> reduction closes the call (see header).
static int force_outer(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *X, const lqcd_fermion *Y, double coef, int accumulate,
                       const lqcd_fermion *dot_with, double *dot_out) {
    if (!ctx->force_buf) {
        const size_t fbytes = (size_t)ctx->g.nblk * 4 * 9 * 32 * sizeof(cplx);
        CUDA_TRY(ctx, cudaMalloc(&ctx->force_buf, fbytes));
        ctx->force_valid = false;
    }
    if (accumulate && !ctx->force_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "force: accumulate requested but no force has been computed yet");
    ForceArgs A;
    A.out = ctx->force_buf; A.X = X->d; A.Y = Y->d; A.gauge = ctx->gauge; A.g = ctx->g; A.kappa = op->kappa;
    A.coef = coef; A.accumulate = accumulate;
    for (int i = 0; i < 4; i++) A.bc[i] = op->bc[i];
    const int bs = 128, grid = (ctx->g.V + bs - 1) / bs;
    const bool multi = ctx->nranks > 1;
    if (multi) {
        LQCD_TRY(comm_force_halo(ctx, &A.fh));
        force_pack_kernel<<<(A.fh.start[4] + bs - 1) / bs, bs, 0, ctx->stream>>>(A, op->kind);
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
        if (op->kind == LQCD_WILSON) wilson_force_kernel<1><<<grid, bs, 0, ctx->stream>>>(A);
        else                         staggered_force_kernel<1><<<grid, bs, 0, ctx->stream>>>(A);
    } else {
        memset(&A.fh, 0, sizeof A.fh);
        if (op->kind == LQCD_WILSON) wilson_force_kernel<0><<<grid, bs, 0, ctx->stream>>>(A);
        else                         staggered_force_kernel<0><<<grid, bs, 0, ctx->stream>>>(A);
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    if (op->kind == LQCD_WILSON && op->csw != 0.0) LQCD_TRY(clover_force_accumulate(ctx, op, X->d, Y->d, coef));      // clover_force.cu
    ctx->force_valid = true;
    // <dot_with, X> (global).  Across ranks this reduction also ORDERS the force slots: it is enqueued after the force
    // kernel and its all-reduce completes only when every rank has got this far, so it runs on every multi-rank call.
    if (dot_out || multi) {
        double d[2];
        LQCD_TRY(lqcd_blas_dot(ctx, dot_with ? dot_with : X, X, d));
        if (dot_out) *dot_out = d[0];
    }
    return comm_check_error(ctx);
}

static int check_force_args(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *a, const lqcd_fermion *b) {
    if (!ctx || !op || !a || !b) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (op->kind != LQCD_WILSON && op->kind != LQCD_STAGGERED) return lqcd_fail(ctx, LQCD_ERR_ARG, "unknown operator kind %d", op->kind);
    if (a->owner != ctx || b->owner != ctx) return lqcd_fail(ctx, LQCD_ERR_ARG, "field belongs to another context");
    if (a->kind != op->kind || b->kind != op->kind) return lqcd_fail(ctx, LQCD_ERR_ARG, "fermion kind does not match the operator");
    if (op->kind == LQCD_WILSON && op->r != 1.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "force implements r = 1 only");
    if (op->kind == LQCD_WILSON && op->csw != 0.0 && ctx->nranks > 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "force: the clover-term derivative is implemented for a single rank (csw must be 0 across ranks)");
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "force requested before lqcd_gauge_upload");
    return LQCD_OK;
}

// gauge_md.cu: X = (DdagD)^-1 eta from a zero guess, Y = D X, UdSfdU left in ctx->force_buf (no host transfer)
int force_for_md(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, double eps, int maxsteps, int *iters) {
    LQCD_TRY(check_force_args(ctx, op, eta, eta));
    lqcd_fermion *X = nullptr, *Y = nullptr;
    LQCD_TRY(get_scratch(ctx, op->kind, SCR_FORCE_X, &X));
    LQCD_TRY(get_scratch(ctx, op->kind, SCR_FORCE_Y, &Y));
    CUDA_TRY(ctx, cudaMemsetAsync(X->d, 0, X->bytes, ctx->stream));
    int it = 0;
    double rs = 0.0;
    LQCD_TRY(lqcd_solve(ctx, op, X, eta, LQCD_SOLVER_CG, LQCD_OP_DDAGD, eps, maxsteps, &it, &rs, nullptr));
    if (iters) *iters = it;
    LQCD_TRY(lqcd_dslash(ctx, op, Y, X, LQCD_OP_D));
    return force_outer(ctx, op, X, Y, 1.0, 0, nullptr, nullptr);
}

// RHMC (staggered Nf not in {4, 8}; README.md:132, test/test_Nf2.toml): S_f = eta^dag r(DdagD) eta with the partial fractions
// r(x) = alpha0 + sum_j alpha[j] / (x + shifts[j]).  X_j = (DdagD + shifts[j])^-1 eta by ONE multi-shift CG, Y_j = D X_j,
// UdSfdU = sum_j alpha[j] force(X_j, Y_j) accumulated in ctx->force_buf.  Scratch: the solver owns slots 0 .. 2 + LQCD_MAX_SHIFTS,
// the force keeps 8 / 9 for the plain action, so the shifted solutions live above both.
static int rational_solutions(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, const double *shifts, int nshift,
                              double eps, int maxsteps, lqcd_fermion **xs, int *iters) {
    if (!shifts || nshift < 1 || nshift > LQCD_MAX_SHIFTS) return lqcd_fail(ctx, LQCD_ERR_ARG, "rational action: nshift must be in [1, %d]", LQCD_MAX_SHIFTS);
    for (int j = 0; j < nshift; j++) LQCD_TRY(get_scratch(ctx, op->kind, SCR_RATIONAL0 + j, &xs[j]));
    int it = 0;
    double rs = 0.0;
    LQCD_TRY(lqcd_multishift_cg(ctx, op, xs, eta, shifts, nshift, eps, maxsteps, &it, &rs));
    if (iters) *iters = it;
    return LQCD_OK;
}

int force_for_md_rational(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, const double *alpha, const double *shifts, int nshift,
                          double eps, int maxsteps, int *iters) {
    LQCD_TRY(check_force_args(ctx, op, eta, eta));
    if (!alpha) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    lqcd_fermion *xs[LQCD_MAX_SHIFTS], *Y = nullptr;
    LQCD_TRY(rational_solutions(ctx, op, eta, shifts, nshift, eps, maxsteps, xs, iters));
    LQCD_TRY(get_scratch(ctx, op->kind, SCR_FORCE_Y, &Y));
    for (int j = 0; j < nshift; j++) {
        LQCD_TRY(lqcd_dslash(ctx, op, Y, xs[j], LQCD_OP_D));
        LQCD_TRY(force_outer(ctx, op, xs[j], Y, alpha[j], j > 0, nullptr, nullptr));
    }
    return LQCD_OK;
}

// calc_UdSfdU! for the RHMC action: the accumulated force copied to the host arrays (link layout, wing ndw)
extern "C" int lqcd_fermion_force_rational(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, const double *alpha, const double *shifts,
                                           int nshift, double eps, int maxsteps, double *const out_mu[4], int ndw, int *iters) {
    if (!out_mu) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (ndw < 0 || ndw > 4) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad wing width %d", ndw);
    LQCD_TRY(check_force_args(ctx, op, eta, eta));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LQCD_TRY(force_for_md_rational(ctx, op, eta, alpha, shifts, nshift, eps, maxsteps, iters));
    return download_links_from(ctx, ctx->force_buf, out_mu, ndw);
}

// y = alpha0 x + sum_j alpha[j] (DdagD + shifts[j])^-1 x: the rational heat bath eta = r_hb(DdagD) xi and, with `dot_out`,
// the action <x, r(DdagD) x> of evaluate_FermiAction, each as one call without host round trips per pole
extern "C" int lqcd_rational_apply(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *x, double alpha0,
                                   const double *alpha, const double *shifts, int nshift, double eps, int maxsteps, int *iters, double *dot_out) {
    if (!ctx || !op || !y || !x || !alpha) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (y == x) return lqcd_fail(ctx, LQCD_ERR_ARG, "rational apply: output aliases input");
    if (x->owner != ctx || y->owner != ctx || x->kind != op->kind || y->kind != op->kind) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad fields");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    lqcd_fermion *xs[LQCD_MAX_SHIFTS];
    LQCD_TRY(rational_solutions(ctx, op, x, shifts, nshift, eps, maxsteps, xs, iters));
    LQCD_TRY(blas_copy(ctx, y->d, x->d, x->bytes / sizeof(cplx)));
    LQCD_TRY(lqcd_blas_scale(ctx, alpha0, 0.0, y));
    for (int j = 0; j < nshift; j++) LQCD_TRY(lqcd_blas_axpy(ctx, alpha[j], 0.0, xs[j], y));
    if (dot_out) {
        double d[2];
        LQCD_TRY(lqcd_blas_dot(ctx, x, y, d));
        *dot_out = d[0];
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return comm_check_error(ctx);
}

extern "C" int lqcd_fermion_force_xy(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *X, const lqcd_fermion *Y, double coef, int accumulate) {
    LQCD_TRY(check_force_args(ctx, op, X, Y));
    if (X == Y) return lqcd_fail(ctx, LQCD_ERR_ARG, "force: X aliases Y");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LQCD_TRY(force_outer(ctx, op, X, Y, coef, accumulate, nullptr, nullptr));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // exported calls are blocking-on-return (include/lqcd_b200.h); internal callers use force_outer
    return comm_check_error(ctx);
}

extern "C" int lqcd_fermion_force_download(lqcd_ctx *ctx, double *const out_mu[4], int ndw) {
    if (!ctx || !out_mu) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (ndw < 0 || ndw > 4) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad wing width %d", ndw);
    if (!ctx->force_buf || !ctx->force_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "no force on the device");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return download_links_from(ctx, ctx->force_buf, out_mu, ndw);
}

extern "C" int lqcd_fermion_force(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, lqcd_fermion *x_inout,
                                  double eps, int maxsteps, double *const out_mu[4], int ndw, int *iters, double *action) {
    if (!out_mu) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (ndw < 0 || ndw > 4) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad wing width %d", ndw);
    LQCD_TRY(check_force_args(ctx, op, eta, eta));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    lqcd_fermion *X = x_inout, *Y = nullptr;
    if (!X) {
        LQCD_TRY(get_scratch(ctx, op->kind, SCR_FORCE_X, &X));
        CUDA_TRY(ctx, cudaMemsetAsync(X->d, 0, X->bytes, ctx->stream));
    }
    LQCD_TRY(get_scratch(ctx, op->kind, SCR_FORCE_Y, &Y));
    int it = 0;
    double rs = 0.0;
    LQCD_TRY(lqcd_solve(ctx, op, X, eta, LQCD_SOLVER_CG, LQCD_OP_DDAGD, eps, maxsteps, &it, &rs, nullptr));
    if (iters) *iters = it;
    LQCD_TRY(lqcd_dslash(ctx, op, Y, X, LQCD_OP_D));
    LQCD_TRY(force_outer(ctx, op, X, Y, 1.0, 0, eta, action));          // S_f = <eta, X>
    return download_links_from(ctx, ctx->force_buf, out_mu, ndw);
}
