// clover_force.cu -- clover-term part of the Wilson-clover pseudofermion force (gather form, deterministic).
//
// New capability like the clover term itself (not reachable from run_LQCD; the reference's disabled test
// test/test_wilsonclover.jl runs HMC with it).  Restated from the oracle (oracle/lqcd_oracle.c: orc_clover_force, pinned by finite
// differences of the action): with X = (M^dag M)^-1 phi, Y = M X,
//     dS = -2 Re[Y^dag dM X],  dM = dA - kappa dH:   hopping part = wilson_force_kernel (force.cu), and
//     clover part  = -2 Re sum_{n,p} (i c / 8) tr[ dQ_p(n) K_p(n) ],   c = kappa csw,  K = Lambda' + Lambda'^dag,
//                    Lambda_p(n)[b,a] = sum sigma_p[al,be] X(n)[be,b] conj(Y(n)[al,a]),  Lambda' = traceless part,
// Q_p(n) the four-leaf sum of clover.cu.  Kernel 1 builds the six K_p(n) per site.  Kernel 2 GATHERS: one thread per link (m, rho)
// visits the 3 planes x 4 leaves x 2 positions where U_rho(m) occurs, finds the leaf's base site n, walks the leaf from n and
// adds  (i c / 8) U S K P  (forward occurrence; P / S = links before / after it) or  -(i c / 8) S K P U^dag  (backward) -- each
// link is written by exactly one thread, so the force is bit-reproducible.  Single rank.
#include "lqcd_internal.cuh"
#include "site_map.cuh"
#include <complex>

struct CloverSigmaF { cplx s[6][2][2][2]; };     // [plane][chirality block][row][col] (clover.cu: make_sigma)
void clover_sigma_tables(cplx (*out)[2][2][2]);  // clover.cu

typedef cplx M3[3][3];

__device__ __forceinline__ void m3_mul(M3 &c, const M3 &a, const M3 &b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            cplx s = cmake(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < 3; k++) cfma(s, a[i][k], b[k][j]);
            r[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c[i][j] = r[i][j];
}
__device__ __forceinline__ void m3_ld(M3 &m, const cplx *p, bool adj) {      // p: element (0,0) of an AoSoA-32 3x3 record
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const cplx v = p[(a * 3 + b) * 32];
            if (adj) m[b][a] = cmake(v.x, -v.y); else m[a][b] = v;
        }
}
__device__ __forceinline__ void m3_one(M3 &m) {
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) m[a][b] = cmake(a == b ? 1.0 : 0.0, 0.0);
}

struct CfArgs {
    cplx *out;               // force buffer, link layout (accumulated into)
    cplx *kf;                // K field: [((blk*6 + p)*9 + e)*32 + lane]
    const cplx *X, *Y, *gauge;
    Geom g;
    double coef;             // (kappa csw / 8) * caller's weight
    CloverSigmaF S;
};

// kernel 1: K_p(n) for the six planes
__global__ void __launch_bounds__(128) clover_k_kernel(const CfArgs A) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.g.V) return;
    const cplx *xp = A.X + (size_t)(s >> 5) * (12 * 32) + (s & 31), *yp = A.Y + (size_t)(s >> 5) * (12 * 32) + (s & 31);
    cplx x[4][3], y[4][3];
#pragma unroll
    for (int al = 0; al < 4; al++)
#pragma unroll
        for (int c = 0; c < 3; c++) { x[al][c] = xp[(3 * al + c) * 32]; y[al][c] = yp[(3 * al + c) * 32]; }
    for (int p = 0; p < 6; p++) {
        M3 L;
        cplx tr = cmake(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int a = 0; a < 3; a++) {
                cplx acc = cmake(0.0, 0.0);
#pragma unroll
                for (int blk = 0; blk < 2; blk++)
#pragma unroll
                    for (int i = 0; i < 2; i++)
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const cplx xy = cmulc(y[2 * blk + i][a], x[2 * blk + j][b]);        // conj(Y[al,a]) X[be,b]
                            cfma(acc, A.S.s[p][blk][i][j], xy);
                        }
                L[b][a] = acc;
            }
#pragma unroll
        for (int a = 0; a < 3; a++) tr = cadd(tr, L[a][a]);
#pragma unroll
        for (int a = 0; a < 3; a++) L[a][a] = cmake(L[a][a].x - tr.x / 3.0, L[a][a].y - tr.y / 3.0);
        cplx *dst = A.kf + ((size_t)(s >> 5) * 6 + p) * (9 * 32) + (s & 31);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) dst[(i * 3 + j) * 32] = cmake(L[i][j].x + L[j][i].x, L[i][j].y - L[j][i].y);      // K = L + L^dag
    }
}

// kernel 2: one thread per link (m, rho)
__global__ void __launch_bounds__(128) clover_force_kernel(const CfArgs A) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.g.V * 4) return;
    const int m = idx % A.g.V, rho = idx / A.g.V;
    const int d[4] = {A.g.X, A.g.Y, A.g.Z, A.g.T};
    int cm[4];
    site_coords(A.g, m, cm[0], cm[1], cm[2], cm[3]);
    M3 acc;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) acc[a][b] = cmake(0.0, 0.0);
    const int sg[4][4] = {{+1, +1, -1, -1}, {+1, -1, -1, +1}, {-1, -1, +1, +1}, {-1, +1, +1, -1}};
    int p = 0;
    for (int mu = 0; mu < 4; mu++)
        for (int nu = mu + 1; nu < 4; nu++, p++) {
            if (mu != rho && nu != rho) continue;
            for (int leaf = 0; leaf < 4; leaf++) {
                const int dir[4] = {(leaf & 1) ? nu : mu, (leaf & 1) ? mu : nu, (leaf & 1) ? nu : mu, (leaf & 1) ? mu : nu};
                for (int k = 0; k < 4; k++) {
                    if (dir[k] != rho) continue;
                    // base site n: the link of step k sits at n + (displacement of steps < k) [- rho if traversed backwards]
                    int cn[4] = {cm[0], cm[1], cm[2], cm[3]};
                    for (int q = 0; q < k; q++) cn[dir[q]] -= sg[leaf][q];
                    if (sg[leaf][k] < 0) cn[rho] += 1;
                    for (int i = 0; i < 4; i++) cn[i] = (cn[i] % d[i] + d[i]) % d[i];
                    const int n = cn[0] + d[0] * (cn[1] + d[1] * (cn[2] + d[2] * cn[3]));
                    // walk the leaf from n: P = L_1 .. L_{k-1}, Lk, S = L_{k+1} .. L_4
                    M3 P, S, Lk, t;
                    m3_one(P); m3_one(S);
                    int c[4] = {cn[0], cn[1], cn[2], cn[3]};
                    for (int q = 0; q < 4; q++) {
                        const int dq = dir[q];
                        M3 L;
                        if (sg[leaf][q] > 0) {
                            const int s = c[0] + d[0] * (c[1] + d[1] * (c[2] + d[2] * c[3]));
                            m3_ld(L, A.gauge + ((size_t)(s >> 5) * 4 + dq) * (9 * 32) + (s & 31), false);
                            c[dq] = (c[dq] + 1) % d[dq];
                        } else {
                            c[dq] = (c[dq] + d[dq] - 1) % d[dq];
                            const int s = c[0] + d[0] * (c[1] + d[1] * (c[2] + d[2] * c[3]));
                            m3_ld(L, A.gauge + ((size_t)(s >> 5) * 4 + dq) * (9 * 32) + (s & 31), true);
                        }
                        if (q < k) m3_mul(P, P, L);
                        else if (q == k) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Lk[i][j] = L[i][j]; }
                        else m3_mul(S, S, L);
                    }
                    M3 K;
                    m3_ld(K, A.kf + ((size_t)(n >> 5) * 6 + p) * (9 * 32) + (n & 31), false);
                    m3_mul(t, S, K);
                    m3_mul(t, t, P);                                   // R = S K P
                    double sign = 1.0;
                    if (sg[leaf][k] > 0) m3_mul(t, Lk, t);             // U R
                    else { m3_mul(t, t, Lk); sign = -1.0; }            // -R U^dag
                    for (int i = 0; i < 3; i++)
                        for (int j = 0; j < 3; j++) { acc[i][j].x += sign * t[i][j].x; acc[i][j].y += sign * t[i][j].y; }
                }
            }
        }
    cplx *o = A.out + ((size_t)(m >> 5) * 4 + rho) * (9 * 32) + (m & 31);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            cplx v = o[(i * 3 + j) * 32];
            v.x += -A.coef * acc[i][j].y;                              // += (i coef) acc
            v.y += A.coef * acc[i][j].x;
            o[(i * 3 + j) * 32] = v;
        }
}

// out (ctx->force_buf, already holding the hopping part) += weight * clover-term force of (X, Y)
int clover_force_accumulate(lqcd_ctx *ctx, const lqcd_op *op, const cplx *X, const cplx *Y, double weight) {
    if (ctx->nranks > 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "force: the clover-term derivative is implemented for a single rank");
    if (!ctx->clover_k) CUDA_TRY(ctx, cudaMalloc(&ctx->clover_k, (size_t)ctx->g.nblk * 54 * 32 * sizeof(cplx)));
    CfArgs A;
    A.out = ctx->force_buf; A.kf = ctx->clover_k; A.X = X; A.Y = Y; A.gauge = ctx->gauge; A.g = ctx->g;
    A.coef = weight * op->kappa * op->csw / 8.0;
    clover_sigma_tables(A.S.s);
    const int bs = 128;
    clover_k_kernel<<<(ctx->g.V + bs - 1) / bs, bs, 0, ctx->stream>>>(A);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    clover_force_kernel<<<(ctx->g.V * 4 + bs - 1) / bs, bs, 0, ctx->stream>>>(A);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}
