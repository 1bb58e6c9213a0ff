// clover.cu -- builds the site-local clover term of the Wilson-clover operator from the device links.
//
//     A(n) = 1 + kappa*csw * sum_{mu<nu} sigma_mu_nu (x) [ i F^_mu_nu(n) ],      sigma_mu_nu = (i/2)[gamma_mu, gamma_nu]
//     F^_mu_nu = (Q_mu_nu - Q_mu_nu^dag)/8 made traceless,   Q_mu_nu = sum of the four plaquette leaves around n
//
// New capability behind lqcd_op.csw (BASELINE.json configs[3] names a Wilson-clover CG).  The surveyed wrapper cannot reach
// Wilson-clover (src/system/universe.jl:106-131; Clover_coefficient parsed at src/system/parameter_structs.jl:125 but never
// forwarded), so the convention is the textbook one and is restated identically in the CPU oracle (oracle/lqcd_oracle.c,
// orc_clover_build); the only in-tree evidence, the four-leaf clover of the dead topological-charge code
// (src/measurements/unusedfiles/measure_topological_charge.jl:299-309 leaf paths, :177-200 traceless anti-Hermitian part),
// is followed for the leaves.  A different upstream normalisation is a rescaling of csw.
//
// Storage (read by wilson_kernel.cuh: clover_apply): in this gamma basis gamma_5 = diag(1,1,-1,-1), sigma_mu_nu is block
// diagonal in chirality, so A(n) is two Hermitian 6x6 blocks = 2 x (6 real + 15 complex) = 72 reals = 576 B/site, AoSoA-32:
//     clover[(blk32*36 + 18*b + e)*32 + lane]   e = 0..2: (A00,A11),(A22,A33),(A44,A55);  e = 3 + i(i-1)/2 + j: A_ij, i > j.
// The term depends on the links only: it is rebuilt when the gauge epoch or kappa*csw changes (once per D(U) rebinding),
// never inside a solve.
//
// Parity on hardware: tests/test_gpu_extended.py::test_clover_term_matches_oracle (1e-13 against the oracle's dense blocks).
#include "lqcd_internal.cuh"
#include "link_view.cuh"
#include <complex>
#include <cstring>

struct CloverSigma { cplx s[6][2][2][2]; };     // [plane][chirality block][row][col]

// acc <- acc * op(m)
__device__ __forceinline__ void mul_right(cplx (&acc)[3][3], const cplx (&m)[3][3], bool adj) {
    cplx r[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            cplx s = cmake(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                cplx b = adj ? cmake(m[j][k].x, -m[j][k].y) : m[k][j];
                cfma(s, acc[i][k], b);
            }
            r[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) acc[i][j] = r[i][j];
}

// one thread per site.  Setup kernel (once per gauge update): clarity over speed, local-memory arrays are fine here.
__global__ void __launch_bounds__(64) clover_build_kernel(cplx *clover, LinkView L, Geom g, double coef, CloverSigma S, int traceless) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.V) return;
    int c0[4];
    { int r = s; c0[0] = r % g.X; r /= g.X; c0[1] = r % g.Y; r /= g.Y; c0[2] = r % g.Z; c0[3] = r / g.Z; }
    double diag[2][6];
    cplx off[2][15];
    for (int b = 0; b < 2; b++) {
        for (int i = 0; i < 6; i++) diag[b][i] = 1.0;
        for (int e = 0; e < 15; e++) off[b][e] = cmake(0.0, 0.0);
    }
    int plane = 0;
    for (int mu = 0; mu < 4; mu++)
        for (int nu = mu + 1; nu < 4; nu++, plane++) {
            // leaves as closed paths from n (dir, sign): same four as measure_topological_charge.jl:299-309
            const int dir[4][4] = {{mu, nu, mu, nu}, {nu, mu, nu, mu}, {mu, nu, mu, nu}, {nu, mu, nu, mu}};
            const int sgn[4][4] = {{+1, +1, -1, -1}, {+1, -1, -1, +1}, {-1, -1, +1, +1}, {-1, +1, +1, -1}};
            cplx Q[3][3];
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Q[a][b] = cmake(0.0, 0.0);
            for (int leaf = 0; leaf < 4; leaf++) {
                cplx acc[3][3], m[3][3];
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) acc[a][b] = cmake(a == b ? 1.0 : 0.0, 0.0);
                int c[4] = {c0[0], c0[1], c0[2], c0[3]};
                for (int k = 0; k < 4; k++) {
                    const int dk = dir[leaf][k];
                    if (sgn[leaf][k] > 0) { fetch_link(m, L, g, c, dk); mul_right(acc, m, false); c[dk] += 1; }
                    else                  { c[dk] -= 1; fetch_link(m, L, g, c, dk); mul_right(acc, m, true); }
                }
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Q[a][b] = cadd(Q[a][b], acc[a][b]);
            }
            // h = i * F^,  F^ = (Q - Q^dag)/8 (traceless): Hermitian 3x3
            cplx F[3][3];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++)
                    F[a][b] = cmake((Q[a][b].x - Q[b][a].x) * 0.125, (Q[a][b].y + Q[b][a].y) * 0.125);
            if (traceless) {
                cplx tr = cadd(cadd(F[0][0], F[1][1]), F[2][2]);
                for (int a = 0; a < 3; a++) F[a][a] = csub(F[a][a], cscale(1.0 / 3.0, tr));
            }
            for (int b = 0; b < 2; b++)
                for (int si = 0; si < 2; si++)
                    for (int sj = 0; sj < 2; sj++) {
                        const cplx sg = S.s[plane][b][si][sj];
                        if (sg.x == 0.0 && sg.y == 0.0) continue;
                        for (int a = 0; a < 3; a++)
                            for (int bb = 0; bb < 3; bb++) {
                                const int i = 3 * si + a, j = 3 * sj + bb;
                                if (i < j) continue;
                                cplx v = cscale(coef, cmul(sg, cmuli(F[a][bb])));
                                if (i == j) diag[b][i] += v.x;
                                else { const int e = i * (i - 1) / 2 + j; off[b][e] = cadd(off[b][e], v); }
                            }
                    }
        }
    cplx *dst = clover + (size_t)(s >> 5) * (36 * 32) + (s & 31);
    for (int b = 0; b < 2; b++) {
        for (int h = 0; h < 3; h++) dst[(18 * b + h) * 32] = cmake(diag[b][2 * h], diag[b][2 * h + 1]);
        for (int e = 0; e < 15; e++) dst[(18 * b + 3 + e) * 32] = off[b][e];
    }
}

// sigma_mu_nu from the library's gamma basis (wilson_spin.cuh / SURVEY.md 8c table), host side
static void make_sigma(CloverSigma &S) {
    typedef std::complex<double> Z;
    const Z I(0.0, 1.0);
    Z g[4][4][4];
    for (auto &m : g) for (auto &r : m) for (auto &e : r) e = 0.0;
    g[0][0][3] = -I;  g[0][1][2] = -I;  g[0][2][1] = I;    g[0][3][0] = I;
    g[1][0][3] = -1.; g[1][1][2] = 1.;  g[1][2][1] = 1.;   g[1][3][0] = -1.;
    g[2][0][2] = -I;  g[2][1][3] = I;   g[2][2][0] = I;    g[2][3][1] = -I;
    g[3][0][2] = -1.; g[3][1][3] = -1.; g[3][2][0] = -1.;  g[3][3][1] = -1.;
    int p = 0;
    for (int mu = 0; mu < 4; mu++)
        for (int nu = mu + 1; nu < 4; nu++, p++)
            for (int b = 0; b < 2; b++)
                for (int i = 0; i < 2; i++)
                    for (int j = 0; j < 2; j++) {
                        Z s = 0.0;
                        for (int k = 0; k < 4; k++) s += g[mu][2 * b + i][k] * g[nu][k][2 * b + j] - g[nu][2 * b + i][k] * g[mu][k][2 * b + j];
                        s *= 0.5 * I;
                        S.s[p][b][i][j] = make_double2(s.real(), s.imag());
                    }
}

void clover_sigma_tables(cplx (*out)[2][2][2]) {            // clover_force.cu
    CloverSigma S;
    make_sigma(S);
    memcpy(out, S.s, sizeof S.s);
}

int comm_link_view(lqcd_ctx *ctx, const cplx **bases);     // comm.cu: every rank's link array (peer mapped); error if unavailable

int ensure_clover(lqcd_ctx *ctx, const lqcd_op *op) {
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "clover term requested before lqcd_gauge_upload");
    const double coef = op->kappa * op->csw;
    if (ctx->clover && ctx->clover_epoch == ctx->gauge_epoch && ctx->clover_coef == coef) return LQCD_OK;
    if (!ctx->clover) CUDA_TRY(ctx, cudaMalloc(&ctx->clover, (size_t)ctx->g.nblk * 36 * 32 * sizeof(cplx)));
    LinkView L;
    memset(&L, 0, sizeof L);
    for (int i = 0; i < 4; i++) { L.pg[i] = ctx->procgrid[i]; L.pc[i] = ctx->pcoord[i]; }
    if (ctx->nranks > 1) LQCD_TRY(comm_link_view(ctx, L.base));
    else L.base[0] = ctx->gauge;
    CloverSigma S;
    make_sigma(S);
    const int bs = 64, grid = (ctx->g.V + bs - 1) / bs;
    clover_build_kernel<<<grid, bs, 0, ctx->stream>>>(ctx->clover, L, ctx->g, coef, S, 1);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->clover_epoch = ctx->gauge_epoch; ctx->clover_coef = coef;
    return LQCD_OK;
}

// C ABI: build (or reuse) the clover term of `op` and, when out != NULL, return it unpacked as dense blocks in the oracle's
// layout: out[(site*2 + b)*36 + i + 6*j] complex (re,im), site = local x-fastest index.  Test / inspection entry point.
extern "C" int lqcd_clover_term(lqcd_ctx *ctx, const lqcd_op *op, double *out) {
    if (!ctx || !op) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (op->kind != LQCD_WILSON) return lqcd_fail(ctx, LQCD_ERR_ARG, "the clover term belongs to the Wilson operator");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LQCD_TRY(ensure_clover(ctx, op));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (!out) return LQCD_OK;
    const size_t n = (size_t)ctx->g.nblk * 36 * 32;
    std::vector<cplx> h(n);
    CUDA_TRY(ctx, cudaMemcpy(h.data(), ctx->clover, n * sizeof(cplx), cudaMemcpyDeviceToHost));
    for (int s = 0; s < ctx->g.V; s++) {
        const cplx *src = h.data() + (size_t)(s >> 5) * (36 * 32) + (s & 31);
        for (int b = 0; b < 2; b++) {
            double *o = out + ((size_t)s * 2 + b) * 72;
            for (int i = 0; i < 6; i++) {
                const cplx d = src[(18 * b + i / 2) * 32];
                o[2 * (i + 6 * i)] = (i & 1) ? d.y : d.x; o[2 * (i + 6 * i) + 1] = 0.0;
                for (int j = 0; j < i; j++) {
                    const cplx c = src[(18 * b + 3 + i * (i - 1) / 2 + j) * 32];
                    o[2 * (i + 6 * j)] = c.x; o[2 * (i + 6 * j) + 1] = c.y;          // A_ij
                    o[2 * (j + 6 * i)] = c.x; o[2 * (j + 6 * i) + 1] = -c.y;         // A_ji = conj
                }
            }
        }
    }
    return LQCD_OK;
}
