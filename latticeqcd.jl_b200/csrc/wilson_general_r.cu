// wilson_general_r.cu -- Wilson operator for r != 1 (the reference forwards params["r"], src/system/universe.jl:115; default 1).
//
//     M = 1 - kappa sum_mu [ (r - g_mu) U_mu(n) T_+mu + (r + g_mu) U_mu^dag(n-mu) T_-mu ]
//       = M_{r=1}  - kappa (r - 1) L,        L x (n) = sum_mu [ U_mu(n) x(n+mu) + U_mu^dag(n-mu) x(n-mu) ]   (spin diagonal)
//
// (r -+ g_mu) is a rank-2 projector only for r = 1, which is what the Dslash kernel exploits.  For any other r the spin-diagonal
// remainder L x is computed by the kernel below into a scratch field and handed to the UNCHANGED r = 1 kernel through its fused
// "y += shift * shift_src" input, so every fused epilogue of the Krylov loops (dot products, norms, the CG residual update) keeps
// working.  M^dag has the same remainder (L is Hermitian).  Costs a second pass over the links: this is the rarely used general
// case, not the hot path.  Single rank; multi-shift CG (which needs the shift input itself), the fermion force and the even-odd
// solve stay r = 1 only.
#include "lqcd_internal.cuh"
#include "site_map.cuh"

struct RTermArgs { cplx *out; const cplx *in; const cplx *gauge; Geom g; double bc[4]; };

template <int MU, int FWD>
__device__ __forceinline__ void lhop(cplx (&acc)[12], const RTermArgs &A, int ns, int ls, double phase) {
    const cplx *sp = A.in + (size_t)(ns >> 5) * (12 * 32) + (ns & 31);
    const cplx *lk = A.gauge + ((size_t)(ls >> 5) * 4 + MU) * (9 * 32) + (ls & 31);
    cplx u[9];
#pragma unroll
    for (int e = 0; e < 9; e++) u[e] = ldg128(lk + e * 32);
#pragma unroll
    for (int sp_i = 0; sp_i < 4; sp_i++) {
        cplx v[3];
#pragma unroll
        for (int c = 0; c < 3; c++) v[c] = cscale(phase, ldg128(sp + (3 * sp_i + c) * 32));
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                if (FWD) cfma(acc[3 * sp_i + a], u[a * 3 + b], v[b]);
                else     cfmac(acc[3 * sp_i + a], u[b * 3 + a], v[b]);
            }
    }
}

__global__ void __launch_bounds__(128) wilson_rterm_kernel(const RTermArgs A) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.g.V) return;
    int x, y, z, t;
    site_coords(A.g, s, x, y, z, t);
    cplx acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = cmake(0.0, 0.0);
#define LPAIR(MU, coord, dim, strd)                                                                      \
    {                                                                                                    \
        const bool wf = (coord == dim - 1), wb = (coord == 0);                                           \
        const int nf = wf ? s - (dim - 1) * (strd) : s + (strd), nb = wb ? s + (dim - 1) * (strd) : s - (strd); \
        lhop<MU, 1>(acc, A, nf, s, wf ? A.bc[MU] : 1.0);                                                 \
        lhop<MU, 0>(acc, A, nb, nb, wb ? A.bc[MU] : 1.0);                                                \
    }
    LPAIR(0, x, A.g.X, 1)
    LPAIR(1, y, A.g.Y, A.g.X)
    LPAIR(2, z, A.g.Z, A.g.X * A.g.Y)
    LPAIR(3, t, A.g.T, A.g.X * A.g.Y * A.g.Z)
#undef LPAIR
    cplx *dst = A.out + (size_t)(s >> 5) * (12 * 32) + (s & 31);
#pragma unroll
    for (int k = 0; k < 12; k++) dst[k * 32] = acc[k];
}

// y = M x (dagger: M^dag x) for r != 1 with the caller's fused epilogue; called from solvers.cu:one_dslash
int wilson_dslash_general_r(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse) {
    if (ctx->nranks > 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "the Wilson operator with r != 1 (r = %g) is implemented for a single rank", op->r);
    if (fuse && fuse->shift_src) return lqcd_fail(ctx, LQCD_ERR_ARG, "multi-shift CG is implemented for the Wilson operator with r = 1 only (got r = %g)", op->r);
    if (fuse && fuse->cta_count > 0) return lqcd_fail(ctx, LQCD_ERR_ARG, "sub-range launch with r != 1");
    lqcd_fermion *z = nullptr;
    LQCD_TRY(get_scratch(ctx, LQCD_WILSON, SCR_RTERM, &z));
    if (z->d == x || z->d == y) return lqcd_fail(ctx, LQCD_ERR_STATE, "r-term scratch field aliases an operand");
    RTermArgs A;
    A.out = z->d; A.in = x; A.gauge = ctx->gauge; A.g = ctx->g;
    for (int i = 0; i < 4; i++) A.bc[i] = op->bc[i];
    const int bs = 128;
    wilson_rterm_kernel<<<(ctx->g.V + bs - 1) / bs, bs, 0, ctx->stream>>>(A);      // (past convergence its consumer is a no-op)
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    lqcd_op op1 = *op;
    op1.r = 1.0;
    DslashFuse f = fuse ? *fuse : DslashFuse();
    f.shift_src = z->d;
    f.shift = -op->kappa * (op->r - 1.0);
    return launch_wilson_dslash(ctx, &op1, y, x, dagger, &f, ctx->stream);
}
