// gauge_md.cu -- gauge-sector molecular dynamics on the device, so that a whole HMC trajectory keeps links, momenta and
// pseudofermions in HBM.
//
// Replaces, for NC = 3 and the plaquette action the wrapper builds (src/system/universe.jl:85-93: beta/2 * (P + P^dag)),
// the steps of src/md/AbstractMD.jl and the integrators of src/md/standardMD.jl:
//     U_update!(U, p, eps, md)          U_mu <- exptU(eps*dtau * p_mu) * U_mu                         AbstractMD.jl:78-99
//     P_update!(U, p, eps, md)          p_mu += (-eps*dtau/NC) * Traceless_antihermitian(U_mu dSdU_mu)  AbstractMD.jl:101-118
//     P_update_fermion!(U, p, eps, md)  p_mu += (-eps*dtau) * Traceless_antihermitian(UdSfdU_mu)        AbstractMD.jl:120-135
//     runMD_QPQ! / runMD_QPQ_sw!        leapfrog, Sexton-Weingarten nesting of the gauge force        standardMD.jl:125-165
//     md.p * md.p / 2, evaluate_GaugeAction                                                           standardHMC.jl:47-50
// Without this the reference-facing path uploads the four link arrays (604 MB at 32^4) and downloads the four force arrays
// every MD step; here the host sees U only at the start and the end of a trajectory.
//
// Momenta live in a link-shaped AoSoA-32 buffer as anti-Hermitian traceless matrices p = sum_a a_a T_a, T_a = i lambda_a/2
// (upstream keeps the eight real a_a [UPSTREAM-RECALL]); p*p/2 = sum a^2/2 = sum_ij |p_ij|^2.  Conventions and factors are
// restated identically in oracle/lqcd_oracle.c (orc_md_*), whose known-answer tests are energy conservation at O(dtau^2)
// and reversibility.  One thread per (site, mu); these kernels are link-bandwidth bound and run a few times per MD step
// next to hundreds of Dslash applications, so they are written for clarity (local 3x3 arrays), not tuned.
// Several ranks: the staples and plaquettes reach one site across the faces (and one corner site, n+mu-nu); those links
// are loaded from the neighbour ranks' peer-mapped link arrays (link_view.cuh, as the clover build does).  Writers and readers
// of the links are ordered across ranks by the in-kernel all-reduce used as a barrier: every rank passes md_barrier after its
// U_update! before anybody evaluates staples, and again after the staples before the next U_update! may overwrite links a
// neighbour is still reading.  The fermion force, CG and momentum updates only touch local data (plus the halo protocols
// of comm.cu / force.cu).
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "rng.cuh"
#include "link_view.cuh"
#include <cstring>

struct MdArgs {
    cplx *gauge;
    cplx *mom;
    const cplx *force;
    Geom g;
    double eps, coef;
    uint64_t seed;
    Reduce red;
    LinkView L;
};

__device__ __forceinline__ const cplx *link_ptr(const cplx *f, int s, int mu) { return f + ((size_t)(s >> 5) * 4 + mu) * (9 * 32) + (s & 31); }
__device__ __forceinline__ void ld3(cplx (&m)[3][3], const cplx *p) {
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) m[a][b] = p[(a * 3 + b) * 32];
}
__device__ __forceinline__ void st3(cplx *p, const cplx (&m)[3][3]) {
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) p[(a * 3 + b) * 32] = m[a][b];
}
// c = op(a) op(b), adjoints selected at compile time
template <int ADJA, int ADJB>
__device__ __forceinline__ void mul3(cplx (&c)[3][3], const cplx (&a)[3][3], const cplx (&b)[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            cplx s = cmake(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const cplx x = ADJA ? cmake(a[k][i].x, -a[k][i].y) : a[i][k];
                const cplx y = ADJB ? cmake(b[j][k].x, -b[j][k].y) : b[k][j];
                cfma(s, x, y);
            }
            c[i][j] = s;
        }
}
__device__ __forceinline__ void ta3(cplx (&a)[3][3], const cplx (&m)[3][3]) {      // traceless anti-Hermitian part
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) a[i][j] = cmake(0.5 * (m[i][j].x - m[j][i].x), 0.5 * (m[i][j].y + m[j][i].y));
    const cplx tr = cscale(1.0 / 3.0, cadd(cadd(a[0][0], a[1][1]), a[2][2]));
#pragma unroll
    for (int i = 0; i < 3; i++) a[i][i] = csub(a[i][i], tr);
}
// exp of a 3x3 matrix: scaling to Frobenius norm <= 1/4, Taylor order 12 (Horner), squaring -- same steps as the oracle's expm3
__device__ __forceinline__ void expm3(cplx (&e)[3][3], const cplx (&x0)[3][3]) {
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) n2 += x0[i][j].x * x0[i][j].x + x0[i][j].y * x0[i][j].y;
    double nrm = sqrt(n2), sc = 1.0;
    int sq = 0;
    while (nrm > 0.25) { nrm *= 0.5; sc *= 0.5; sq++; }
    cplx x[3][3], t[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { x[i][j] = cscale(sc, x0[i][j]); e[i][j] = cmake(i == j ? 1.0 : 0.0, 0.0); }
    for (int k = 12; k >= 1; k--) {                      // e = 1 + (x/k) e
        mul3<0, 0>(t, x, e);
        const double r = 1.0 / (double)k;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) e[i][j] = cmake((i == j ? 1.0 : 0.0) + t[i][j].x * r, t[i][j].y * r);
    }
    for (int q = 0; q < sq; q++) {
        mul3<0, 0>(t, e, e);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) e[i][j] = t[i][j];
    }
}

__device__ __forceinline__ void coords4(const Geom &g, int s, int (&c)[4]) {
    c[0] = s % g.X; s /= g.X; c[1] = s % g.Y; s /= g.Y; c[2] = s % g.Z; c[3] = s / g.Z;
}
// P_update!: p_mu(n) -= eps * beta/(2 NC) * TA( U_mu(n) * sum of the six staples ).  Coordinates may leave the local lattice
// by one step (two directions at once for the corner n+mu-nu): fetch_link maps them to the owning rank.
__global__ void __launch_bounds__(128) md_update_p_gauge_kernel(const MdArgs A) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.g.V * 4) return;
    const int s = idx % A.g.V, mu = idx / A.g.V;
    int c[4];
    coords4(A.g, s, c);
    cplx V[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = cmake(0.0, 0.0);
    for (int nu = 0; nu < 4; nu++) {
        if (nu == mu) continue;
        int cmu[4] = {c[0], c[1], c[2], c[3]}, cnu[4] = {c[0], c[1], c[2], c[3]}, cdn[4] = {c[0], c[1], c[2], c[3]}, cmd[4] = {c[0], c[1], c[2], c[3]};
        cmu[mu] += 1; cnu[nu] += 1; cdn[nu] -= 1; cmd[mu] += 1; cmd[nu] -= 1;
        cplx a[3][3], b[3][3], t[3][3], r[3][3];
        fetch_link(a, A.L, A.g, cmu, nu); fetch_link(b, A.L, A.g, cnu, mu);
        mul3<0, 1>(t, a, b);                                          // U_nu(n+mu) U_mu(n+nu)^dag
        ld3(a, link_ptr(A.gauge, s, nu));
        mul3<0, 1>(r, t, a);                                          // ... U_nu(n)^dag
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = cadd(V[i][j], r[i][j]);
        fetch_link(a, A.L, A.g, cmd, nu); fetch_link(b, A.L, A.g, cdn, mu);
        mul3<1, 1>(t, a, b);                                          // U_nu(n+mu-nu)^dag U_mu(n-nu)^dag
        fetch_link(a, A.L, A.g, cdn, nu);
        mul3<0, 0>(r, t, a);                                          // ... U_nu(n-nu)
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = cadd(V[i][j], r[i][j]);
    }
    cplx U[3][3], M[3][3], T[3][3], P[3][3];
    ld3(U, link_ptr(A.gauge, s, mu));
    mul3<0, 0>(M, U, V);
    ta3(T, M);
    cplx *pp = const_cast<cplx *>(link_ptr(A.mom, s, mu));
    ld3(P, pp);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) P[i][j] = cmake(fma(-A.coef, T[i][j].x, P[i][j].x), fma(-A.coef, T[i][j].y, P[i][j].y));
    st3(pp, P);
}

// sum over the local sites of the six plaquettes Re tr U_mu(n) U_nu(n+mu) U_mu(n+nu)^dag U_nu(n)^dag (all-reduced across ranks)
__global__ void __launch_bounds__(128) md_plaquette_kernel(const MdArgs A) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double red[1] = {0.0};
    if (s < A.g.V) {
        int c[4];
        coords4(A.g, s, c);
        for (int mu = 0; mu < 4; mu++)
            for (int nu = mu + 1; nu < 4; nu++) {
                int cmu[4] = {c[0], c[1], c[2], c[3]}, cnu[4] = {c[0], c[1], c[2], c[3]};
                cmu[mu] += 1; cnu[nu] += 1;
                cplx a[3][3], b[3][3], t[3][3], r[3][3];
                ld3(a, link_ptr(A.gauge, s, mu)); fetch_link(b, A.L, A.g, cmu, nu);
                mul3<0, 0>(t, a, b);
                fetch_link(a, A.L, A.g, cnu, mu);
                mul3<0, 1>(r, t, a);
                ld3(b, link_ptr(A.gauge, s, nu));
                for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) red[0] += r[i][k].x * b[i][k].x + r[i][k].y * b[i][k].y;   // Re tr(r b^dag)
            }
    }
    grid_reduce_finish<1>(red, A.red, FIN_STORE);
}

// cross-rank barrier: the in-kernel all-reduce of reduce.cuh completes only when every rank has arrived
__global__ void md_barrier_kernel(const MdArgs A) {
    double red[1] = {0.0};
    grid_reduce_finish<1>(red, A.red, FIN_STORE);
}

// P_update_fermion!: p_mu(n) -= eps * TA( UdSfdU_mu(n) ), the force field left on the device by force.cu
__global__ void __launch_bounds__(128) md_update_p_force_kernel(const MdArgs A) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.g.V * 4) return;
    const int s = idx % A.g.V, mu = idx / A.g.V;
    cplx F[3][3], T[3][3], P[3][3];
    ld3(F, link_ptr(A.force, s, mu));
    ta3(T, F);
    cplx *pp = const_cast<cplx *>(link_ptr(A.mom, s, mu));
    ld3(P, pp);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) P[i][j] = cmake(fma(-A.coef, T[i][j].x, P[i][j].x), fma(-A.coef, T[i][j].y, P[i][j].y));
    st3(pp, P);
}

// U_update!: U_mu(n) <- exp(eps p_mu(n)) U_mu(n)
__global__ void __launch_bounds__(128) md_update_u_kernel(const MdArgs A) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.g.V * 4) return;
    const int s = idx % A.g.V, mu = idx / A.g.V;
    cplx P[3][3], E[3][3], U[3][3], W[3][3];
    ld3(P, link_ptr(A.mom, s, mu));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) P[i][j] = cscale(A.coef, P[i][j]);
    expm3(E, P);
    cplx *up = const_cast<cplx *>(link_ptr(A.gauge, s, mu));
    ld3(U, up);
    mul3<0, 0>(W, E, U);
    st3(up, W);
}

// p*p/2 = sum |p_ij|^2
__global__ void __launch_bounds__(256) md_kinetic_kernel(const MdArgs A, size_t n) {
    double red[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const cplx v = A.mom[i];
        red[0] = fma(v.x, v.x, red[0]); red[0] = fma(v.y, v.y, red[0]);
    }
    grid_reduce_finish<1>(red, A.red, FIN_STORE);
}

// gauss_distribution!(p): a_a ~ N(0,1), p = (i/2) sum_a a_a lambda_a; counter = global link index (decomposition independent)
__global__ void __launch_bounds__(128) md_momenta_kernel(const MdArgs A) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.g.V * 4) return;
    const int s = idx % A.g.V, mu = idx / A.g.V;
    const uint64_t link = (uint64_t)global_site(A.g, s) * 4 + mu;
    double a[8];
    for (int k = 0; k < 4; k++) gauss_pair(A.seed, link * 4 + k, a[2 * k], a[2 * k + 1]);
    const double r3 = 0.57735026918962576451;      // 1/sqrt(3)
    cplx P[3][3];
    P[0][0] = cmake(0.0, 0.5 * (a[2] + a[7] * r3));
    P[1][1] = cmake(0.0, 0.5 * (-a[2] + a[7] * r3));
    P[2][2] = cmake(0.0, -a[7] * r3);
    P[0][1] = cmake(0.5 * a[1], 0.5 * a[0]);  P[1][0] = cmake(-0.5 * a[1], 0.5 * a[0]);
    P[0][2] = cmake(0.5 * a[4], 0.5 * a[3]);  P[2][0] = cmake(-0.5 * a[4], 0.5 * a[3]);
    P[1][2] = cmake(0.5 * a[6], 0.5 * a[5]);  P[2][1] = cmake(-0.5 * a[6], 0.5 * a[5]);
    st3(const_cast<cplx *>(link_ptr(A.mom, s, mu)), P);
}

// ---- host side -------------------------------------------------------------------------------------------------------------
int upload_links_to(lqcd_ctx *ctx, cplx *dev_links, const double *const U_mu[4], int ndw);          // context.cu
int download_links_from(lqcd_ctx *ctx, const cplx *dev_links, double *const U_mu[4], int ndw);      // context.cu
int comm_check_error(lqcd_ctx *ctx);                                                                // comm.cu
int force_for_md(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, double eps, int maxsteps, int *iters);   // force.cu
int force_for_md_rational(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, const double *alpha, const double *shifts, int nshift,
                          double eps, int maxsteps, int *iters);                                                       // force.cu

// pseudofermion action inside a trajectory: plain eta^dag (DdagD)^-1 eta (nshift = 0) or the RHMC partial fractions
struct MdFermion {
    const lqcd_op *op; const lqcd_fermion *eta;
    const double *alpha, *shifts; int nshift;
    double cg_eps; int cg_maxsteps;
};

static int md_ready(lqcd_ctx *ctx, bool need_mom) {
    if (!ctx) return lqcd_fail(nullptr, LQCD_ERR_ARG, "null ctx");
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "no gauge field on the device");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->mom) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->mom, (size_t)ctx->g.nblk * 4 * 9 * 32 * sizeof(cplx)));
        ctx->mom_valid = false;
    }
    if (need_mom && !ctx->mom_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "no momenta on the device (lqcd_md_momenta_gaussian / _upload first)");
    return LQCD_OK;
}
static MdArgs md_args(lqcd_ctx *ctx) {
    MdArgs A;
    memset(&A, 0, sizeof A);
    A.gauge = ctx->gauge; A.mom = ctx->mom; A.force = ctx->force_buf; A.g = ctx->g; A.red = ctx->red;
    return A;
}
static int md_barrier(lqcd_ctx *ctx) {
    if (ctx->nranks == 1) return LQCD_OK;
    MdArgs A = md_args(ctx);
    md_barrier_kernel<<<1, 32, 0, ctx->stream>>>(A);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}
#define MD_LAUNCH(kernel, A)                                                                  \
    do {                                                                                      \
        const int n_ = ctx->g.V * 4, bs_ = 128;                                               \
        kernel<<<(n_ + bs_ - 1) / bs_, bs_, 0, ctx->stream>>>(A);                             \
        ctx->launches++;                                                                      \
        CUDA_TRY(ctx, cudaGetLastError());                                                    \
    } while (0)

static int md_update_u(lqcd_ctx *ctx, double eps) {
    MdArgs A = md_args(ctx);
    A.coef = eps;
    MD_LAUNCH(md_update_u_kernel, A);
    ctx->gauge_epoch++;                        // clover term / even-odd link copies are stale now
    return LQCD_OK;
}
static int md_update_p_gauge(lqcd_ctx *ctx, double eps, double beta) {
    MdArgs A = md_args(ctx);
    LQCD_TRY(make_link_view(ctx, A.L));
    A.coef = eps * beta / 6.0;                 // beta / (2 NC)
    LQCD_TRY(md_barrier(ctx));                 // every rank has finished writing its links ...
    MD_LAUNCH(md_update_p_gauge_kernel, A);
    LQCD_TRY(md_barrier(ctx));                 // ... and nobody overwrites them while a neighbour still reads
    return LQCD_OK;
}
static int md_update_p_fermion(lqcd_ctx *ctx, const MdFermion &f, double eps, int *iters) {
    // leaves UdSfdU in ctx->force_buf
    if (f.nshift > 0) LQCD_TRY(force_for_md_rational(ctx, f.op, f.eta, f.alpha, f.shifts, f.nshift, f.cg_eps, f.cg_maxsteps, iters));
    else              LQCD_TRY(force_for_md(ctx, f.op, f.eta, f.cg_eps, f.cg_maxsteps, iters));
    MdArgs A = md_args(ctx);
    A.coef = eps;
    MD_LAUNCH(md_update_p_force_kernel, A);
    return LQCD_OK;
}
// runMD_QPQ! (nsw = 0) / runMD_QPQ_sw! (standardMD.jl:125-165); f = nullptr: quenched
static int md_trajectory(lqcd_ctx *ctx, const MdFermion *f, double beta, double dtau, int mdsteps, int nsw, long long *cg_iters_total) {
    if (mdsteps < 1 || nsw < 0 || (nsw & 1) || !(dtau > 0.0)) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad MD parameters (mdsteps >= 1, nsw even >= 0, dtau > 0)");
    if (f && !f->eta) return lqcd_fail(ctx, LQCD_ERR_ARG, "dynamical run needs the pseudofermion field eta");
    LQCD_TRY(md_ready(ctx, true));
    long long its = 0;
    for (int step_i = 0; step_i < mdsteps; step_i++) {
        int it = 0;
        if (nsw == 0) {
            LQCD_TRY(md_update_u(ctx, 0.5 * dtau));
            LQCD_TRY(md_update_p_gauge(ctx, dtau, beta));
            if (f) { LQCD_TRY(md_update_p_fermion(ctx, *f, dtau, &it)); its += it; }
            LQCD_TRY(md_update_u(ctx, 0.5 * dtau));
        } else {
            for (int half = 0; half < 2; half++) {
                for (int isw = 0; isw < nsw / 2; isw++) {
                    LQCD_TRY(md_update_u(ctx, 0.5 * dtau / nsw));
                    LQCD_TRY(md_update_p_gauge(ctx, dtau / nsw, beta));
                    LQCD_TRY(md_update_u(ctx, 0.5 * dtau / nsw));
                }
                if (half == 0 && f) { LQCD_TRY(md_update_p_fermion(ctx, *f, dtau, &it)); its += it; }
            }
        }
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (cg_iters_total) *cg_iters_total = its;
    return LQCD_OK;
}

extern "C" int lqcd_md_momenta_gaussian(lqcd_ctx *ctx, uint64_t seed) {
    LQCD_TRY(md_ready(ctx, false));
    MdArgs A = md_args(ctx);
    A.seed = seed;
    MD_LAUNCH(md_momenta_kernel, A);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->mom_valid = true;
    return LQCD_OK;
}
extern "C" int lqcd_md_momenta_upload(lqcd_ctx *ctx, const double *const P_mu[4], int ndw) {
    if (!P_mu) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    LQCD_TRY(md_ready(ctx, false));
    LQCD_TRY(upload_links_to(ctx, ctx->mom, P_mu, ndw));
    ctx->mom_valid = true;
    return LQCD_OK;
}
extern "C" int lqcd_md_momenta_download(lqcd_ctx *ctx, double *const P_mu[4], int ndw) {
    if (!P_mu) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    LQCD_TRY(md_ready(ctx, true));
    return download_links_from(ctx, ctx->mom, P_mu, ndw);
}
extern "C" int lqcd_md_kinetic(lqcd_ctx *ctx, double *out) {
    if (!out) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    LQCD_TRY(md_ready(ctx, true));
    MdArgs A = md_args(ctx);
    const size_t n = (size_t)ctx->g.nblk * 4 * 9 * 32;
    md_kinetic_kernel<<<reduce_grid(ctx), 256, 0, ctx->stream>>>(A, n);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st_host, ctx->red.st, sizeof(SolverState), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = ctx->st_host->red[0];
    return LQCD_OK;
}
extern "C" int lqcd_md_gauge_action(lqcd_ctx *ctx, double beta, double *out) {
    if (!out) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    LQCD_TRY(md_ready(ctx, false));
    MdArgs A = md_args(ctx);
    LQCD_TRY(make_link_view(ctx, A.L));
    LQCD_TRY(md_barrier(ctx));
    const int bs = 128;
    md_plaquette_kernel<<<(ctx->g.V + bs - 1) / bs, bs, 0, ctx->stream>>>(A);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st_host, ctx->red.st, sizeof(SolverState), cudaMemcpyDeviceToHost, ctx->stream));
    LQCD_TRY(md_barrier(ctx));                 // (after the copy: the barrier's own reduction overwrites st->red)
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    LQCD_TRY(comm_check_error(ctx));
    *out = -(beta / 3.0) * ctx->st_host->red[0];                      // -(beta/NC) sum_plaq Re tr U_p (global)
    return LQCD_OK;
}
extern "C" int lqcd_md_update_u(lqcd_ctx *ctx, double eps) {
    LQCD_TRY(md_ready(ctx, true));
    LQCD_TRY(md_update_u(ctx, eps));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}
extern "C" int lqcd_md_update_p(lqcd_ctx *ctx, double eps, double beta) {
    LQCD_TRY(md_ready(ctx, true));
    LQCD_TRY(md_update_p_gauge(ctx, eps, beta));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}
extern "C" int lqcd_md_update_p_fermion(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, double eps, double cg_eps, int cg_maxsteps, int *iters) {
    if (!op || !eta) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    LQCD_TRY(md_ready(ctx, true));
    const MdFermion f = {op, eta, nullptr, nullptr, 0, cg_eps, cg_maxsteps};
    LQCD_TRY(md_update_p_fermion(ctx, f, eps, iters));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}
// the same for the RHMC action eta^dag [alpha0 + sum_j alpha[j] / (DdagD + shifts[j])] eta (alpha0 does not move the links)
extern "C" int lqcd_md_update_p_fermion_rational(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, const double *alpha, const double *shifts,
                                                 int nshift, double eps, double cg_eps, int cg_maxsteps, int *iters) {
    if (!op || !eta || !alpha || !shifts || nshift < 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument / nshift < 1");
    LQCD_TRY(md_ready(ctx, true));
    const MdFermion f = {op, eta, alpha, shifts, nshift, cg_eps, cg_maxsteps};
    LQCD_TRY(md_update_p_fermion(ctx, f, eps, iters));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

// runMD!(U, md): MDsteps leapfrog steps of size dtau on the device.  nsw = 0: runMD_QPQ! (standardMD.jl:125-141);
// nsw > 0 (even): runMD_QPQ_sw!, the gauge force integrated with nsw sub-steps around one fermion force (standardMD.jl:143-165).
// op == NULL: quenched.  cg_iters_total (nullable) accumulates the CG iterations of all fermion-force solves.
extern "C" int lqcd_md_trajectory(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, double beta, double dtau, int mdsteps, int nsw,
                                  double cg_eps, int cg_maxsteps, long long *cg_iters_total) {
    if (op && !eta) return lqcd_fail(ctx, LQCD_ERR_ARG, "dynamical run needs the pseudofermion field eta");
    const MdFermion f = {op, eta, nullptr, nullptr, 0, cg_eps, cg_maxsteps};
    return md_trajectory(ctx, op ? &f : nullptr, beta, dtau, mdsteps, nsw, cg_iters_total);
}
// RHMC trajectory (test/test_Nf2.toml, BASELINE config 5): the fermion force of every step is ONE multi-shift CG + nshift outer products
extern "C" int lqcd_md_trajectory_rational(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, const double *alpha, const double *shifts,
                                           int nshift, double beta, double dtau, int mdsteps, int nsw, double cg_eps, int cg_maxsteps,
                                           long long *cg_iters_total) {
    if (!op || !eta || !alpha || !shifts || nshift < 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument / nshift < 1");
    const MdFermion f = {op, eta, alpha, shifts, nshift, cg_eps, cg_maxsteps};
    return md_trajectory(ctx, &f, beta, dtau, mdsteps, nsw, cg_iters_total);
}
