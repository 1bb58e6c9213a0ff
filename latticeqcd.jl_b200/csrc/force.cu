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
#include "lqcd_internal.cuh"
#include "wilson_spin.cuh"
#include "site_map.cuh"

struct ForceArgs {
    cplx *out;               // device link layout [((blk*4+mu)*9 + a*3+b)*32 + lane]
    const cplx *X, *Y, *gauge;
    Geom g;
    double kappa;
    double bc[4];
};

__device__ __forceinline__ void load_spinor(cplx (&p)[12], const cplx *f, int s) {
    const cplx *sp = f + (size_t)(s >> 5) * (12 * 32) + (s & 31);
#pragma unroll
    for (int k = 0; k < 12; k++) p[k] = ldg128(sp + k * 32);
}

template <int MU>
__device__ __forceinline__ void wilson_force_dir(const ForceArgs &A, int s, int coord, int dim, int stride,
                                                 const cplx (&Xn)[12], const cplx (&Yn)[12]) {
    const bool w = (coord == dim - 1);
    const int ns = w ? s - (dim - 1) * stride : s + stride;
    const double phase = w ? A.bc[MU] : 1.0;
    cplx Xf[12], Yf[12];
    load_spinor(Xf, A.X, ns);
    load_spinor(Yf, A.Y, ns);
    cplx hx0[3], hx1[3], hy0[3], hy1[3], px0[3], px1[3], py0[3], py1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        project<MU, -1>(hx0[c], hx1[c], Xf[c], Xf[3 + c], Xf[6 + c], Xf[9 + c]);      // P- X(n+mu)
        project<MU, +1>(hy0[c], hy1[c], Yf[c], Yf[3 + c], Yf[6 + c], Yf[9 + c]);      // P+ Y(n+mu)
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
            o[(a * 3 + b) * 32] = cmake(A.kappa * (t2.x - t1.x), A.kappa * (t2.y - t1.y));
        }
}

__global__ void __launch_bounds__(128) wilson_force_kernel(const ForceArgs A) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.g.V) return;
    int x, y, z, t;
    site_coords(A.g, s, x, y, z, t);
    cplx Xn[12], Yn[12];
    load_spinor(Xn, A.X, s);
    load_spinor(Yn, A.Y, s);
    wilson_force_dir<0>(A, s, x, A.g.X, 1, Xn, Yn);
    wilson_force_dir<1>(A, s, y, A.g.Y, A.g.X, Xn, Yn);
    wilson_force_dir<2>(A, s, z, A.g.Z, A.g.X * A.g.Y, Xn, Yn);
    wilson_force_dir<3>(A, s, t, A.g.T, A.g.X * A.g.Y * A.g.Z, Xn, Yn);
}

template <int MU>
__device__ __forceinline__ void stag_force_dir(const ForceArgs &A, int s, int coord, int dim, int stride, double eta,
                                               const cplx (&Xn)[3], const cplx (&Yn)[3]) {
    const bool w = (coord == dim - 1);
    const int ns = w ? s - (dim - 1) * stride : s + stride;
    const double phase = w ? A.bc[MU] : 1.0;
    const cplx *xp = A.X + (size_t)(ns >> 5) * (3 * 32) + (ns & 31), *yp = A.Y + (size_t)(ns >> 5) * (3 * 32) + (ns & 31);
    cplx xf[3], yf[3], hx[3], hy[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { xf[c] = cscale(phase, ldg128(xp + c * 32)); yf[c] = cscale(phase, ldg128(yp + c * 32)); }
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
    const double cf = 0.5 * eta;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            cplx t = cmake(0, 0);
            cfmac(t, Yn[b], hx[a]);
            cfmac(t, hy[b], Xn[a]);
            o[(a * 3 + b) * 32] = cscale(cf, t);
        }
}

__global__ void __launch_bounds__(128) staggered_force_kernel(const ForceArgs A) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.g.V) return;
    int x, y, z, t;
    site_coords(A.g, s, x, y, z, t);
    cplx Xn[3], Yn[3];
    const cplx *xp = A.X + (size_t)(s >> 5) * (3 * 32) + (s & 31), *yp = A.Y + (size_t)(s >> 5) * (3 * 32) + (s & 31);
#pragma unroll
    for (int c = 0; c < 3; c++) { Xn[c] = ldg128(xp + c * 32); Yn[c] = ldg128(yp + c * 32); }
    const int gx = x + A.g.o[0], gy = y + A.g.o[1], gz = z + A.g.o[2];
    stag_force_dir<0>(A, s, x, A.g.X, 1, 1.0, Xn, Yn);
    stag_force_dir<1>(A, s, y, A.g.Y, A.g.X, (gx & 1) ? -1.0 : 1.0, Xn, Yn);
    stag_force_dir<2>(A, s, z, A.g.Z, A.g.X * A.g.Y, ((gx + gy) & 1) ? -1.0 : 1.0, Xn, Yn);
    stag_force_dir<3>(A, s, t, A.g.T, A.g.X * A.g.Y * A.g.Z, ((gx + gy + gz) & 1) ? -1.0 : 1.0, Xn, Yn);
}

int download_links_from(lqcd_ctx *ctx, const cplx *dev_links, double *const U_mu[4], int ndw);     // context.cu

extern "C" int lqcd_fermion_force(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, lqcd_fermion *x_inout,
                                  double eps, int maxsteps, double *const out_mu[4], int ndw, int *iters, double *action) {
    if (!ctx || !op || !eta || !out_mu) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (ndw < 0 || ndw > 4) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad wing width %d", ndw);
    if (ctx->nranks > 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "lqcd_fermion_force: single-rank only in this round");
    if (eta->kind != op->kind) return lqcd_fail(ctx, LQCD_ERR_ARG, "fermion kind does not match the operator");
    if (op->kind == LQCD_WILSON && op->r != 1.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "force implements r = 1 only");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    lqcd_fermion *X = x_inout, *Y = nullptr;
    if (!X) {
        LQCD_TRY(get_scratch(ctx, op->kind, 8, &X));
        CUDA_TRY(ctx, cudaMemsetAsync(X->d, 0, X->bytes, ctx->stream));
    }
    LQCD_TRY(get_scratch(ctx, op->kind, 9, &Y));
    int it = 0;
    double rs = 0.0;
    LQCD_TRY(lqcd_solve(ctx, op, X, eta, LQCD_SOLVER_CG, LQCD_OP_DDAGD, eps, maxsteps, &it, &rs, nullptr));
    if (iters) *iters = it;
    LQCD_TRY(lqcd_dslash(ctx, op, Y, X, LQCD_OP_D));
    if (action) {
        double d[2];
        LQCD_TRY(lqcd_blas_dot(ctx, eta, X, d));
        *action = d[0];
    }
    // force field in the device link layout, then the same conversion path as lqcd_gauge_download
    if (!ctx->force_buf) {
        const size_t fbytes = (size_t)ctx->g.nblk * 4 * 9 * 32 * sizeof(cplx);
        CUDA_TRY(ctx, cudaMalloc(&ctx->force_buf, fbytes));
    }
    cplx *fbuf = ctx->force_buf;
    ForceArgs A;
    A.out = fbuf; A.X = X->d; A.Y = Y->d; A.gauge = ctx->gauge; A.g = ctx->g; A.kappa = op->kappa;
    for (int i = 0; i < 4; i++) A.bc[i] = op->bc[i];
    const int bs = 128, grid = (ctx->g.V + bs - 1) / bs;
    if (op->kind == LQCD_WILSON) wilson_force_kernel<<<grid, bs, 0, ctx->stream>>>(A);
    else                         staggered_force_kernel<<<grid, bs, 0, ctx->stream>>>(A);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    int rc = LQCD_OK;
    if (e != cudaSuccess) rc = lqcd_fail(ctx, LQCD_ERR_CUDA, "force kernel -> %s", cudaGetErrorString(e));
    if (rc == LQCD_OK) rc = download_links_from(ctx, fbuf, out_mu, ndw);
    return rc;
}
