// staggered_eo.cu -- CG for the staggered even-site system on checkerboarded half fields.
//
// Staggered Nf = 4 is the reference's default staggered setup (test/test_staggered.toml, "Staggered SU(3) with 4 tastes",
// test/runtests.jl:101-112): the pseudofermion lives on even sites only (lqcd_b200/api.py FermiActionB200).
// With D = m + Dh, Dh anti-Hermitian and connecting opposite parities, D^dag D = m^2 - Dh^2 does not couple the parities, so for
// an even-site source the system every HMC force / action evaluation solves is
//
//     A_ee x_e = b_e,      A_ee = m^2 - Dh_eo Dh_oe   (Hermitian, >= m^2),
//
// which the full-lattice CG solves with half of its sites carrying exact zeros.  Here the same CG (solvers.cu: solve_impl, all
// scalars device resident) runs on half fields: per iteration two parity hops over V/2 sites each -- 2 x (1152 B links + 48 B in
// + 48 B out) per even site = 1248 B per full-lattice site against 2 x 672 = 1344 B, and, more to the point, vector updates and
// reductions of half the length and fields that fit L2 twice as long.  The arithmetic per even site is that of the full-lattice
// solve (the odd half only ever contributed zeros), so iteration counts agree up to the rounding of the reduction order.
//
//     first half   (dagger = 0)   t_o = Dh_oe p_e                      reduces |t_o|^2 + m^2 |p_e|^2 = <p, A_ee p>   (norm form)
//     second half  (dagger = 1)   q_e = m^2 p_e - Dh_eo t_o            with the solver's fused epilogue (r -= alpha q, |r|^2, <w, q>)
//
// Layout and link split: wilson_eo.cu (eo_common.cuh).  Single rank.  Parity on hardware: tests/test_gpu_extended.py.
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "site_map.cuh"
#include "eo_common.cuh"
#include <cstring>

int solve_impl(lqcd_ctx *ctx, const lqcd_op *op, cplx *x, const cplx *bb, size_t n, int method, int target,
               double eps, int maxsteps, int *iters, double *resid_sq, double *hist);      // solvers.cu

struct StagEoArgs {
    cplx *out;               // output half field (parity `parity`)
    const cplx *in;          // input half field (opposite parity)
    const cplx *xsrc;        // nullable: out = xcoef * xsrc + coef * Dh in   (same parity as out)
    const cplx *norm_src;    // nullable: red2 += norm_coef * |norm_src|^2 at the thread's own half index
    double coef, xcoef, norm_coef;
    const cplx *g_out, *g_in;
    Geom gh;
    int parity;
    double bc[4];
    DslashFuse fuse;
    Reduce red;
};

template <int MU, int FWD>
__device__ __forceinline__ void shop_h(cplx (&acc)[3], const cplx *__restrict__ in, const cplx *__restrict__ gauge, int ns, int ls, double coef) {
    const cplx *sp = in + (size_t)(ns >> 5) * (3 * 32) + (ns & 31);
    cplx v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) v[c] = cscale(coef, ldg128(sp + c * 32));
    const cplx *lk = gauge + ((size_t)(ls >> 5) * 4 + MU) * (9 * 32) + (ls & 31);
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = 0; b < 3; b++) {
            if (FWD) cfma(acc[a], ldg128(lk + (a * 3 + b) * 32), v[b]);
            else     cfmac(acc[a], ldg128(lk + (b * 3 + a) * 32), v[b]);
        }
    }
}

// out(n) = [xcoef * xsrc(n)] + coef * sum_mu (eta_mu(n) / 2) [ U_mu(n) in(n+mu) - U_mu^dag(n-mu) in(n-mu) ],  n of parity `parity`
__global__ void __launch_bounds__(256) staggered_eo_hop_kernel(const StagEoArgs A) {
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
        const int x = 2 * xh + odd_row;
        const double e1 = (x & 1) ? -1.0 : 1.0, e2 = ((x + y) & 1) ? -1.0 : 1.0, e3 = ((x + y + z) & 1) ? -1.0 : 1.0;
        cplx acc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) acc[k] = cmake(0.0, 0.0);
        {   // x direction (eta_1 = 1): the neighbour of the other parity sits at the same xh or one step away
            const bool wf = odd_row && (xh == Xh - 1);           // x == X-1
            const int nf = odd_row ? (wf ? h - (Xh - 1) : h + 1) : h;
            shop_h<0, 1>(acc, A.in, A.g_out, nf, h, 0.5 * (wf ? A.bc[0] : 1.0));
            const bool wb = !odd_row && (xh == 0);               // x == 0
            const int nb = odd_row ? h : (wb ? h + (Xh - 1) : h - 1);
            shop_h<0, 0>(acc, A.in, A.g_in, nb, nb, -0.5 * (wb ? A.bc[0] : 1.0));
        }
        {
            const int st = Xh;
            const bool wf = (y == A.gh.Y - 1), wb = (y == 0);
            const int nf = wf ? h - (A.gh.Y - 1) * st : h + st, nb = wb ? h + (A.gh.Y - 1) * st : h - st;
            shop_h<1, 1>(acc, A.in, A.g_out, nf, h, 0.5 * e1 * (wf ? A.bc[1] : 1.0));
            shop_h<1, 0>(acc, A.in, A.g_in, nb, nb, -0.5 * e1 * (wb ? A.bc[1] : 1.0));
        }
        {
            const int st = Xh * A.gh.Y;
            const bool wf = (z == A.gh.Z - 1), wb = (z == 0);
            const int nf = wf ? h - (A.gh.Z - 1) * st : h + st, nb = wb ? h + (A.gh.Z - 1) * st : h - st;
            shop_h<2, 1>(acc, A.in, A.g_out, nf, h, 0.5 * e2 * (wf ? A.bc[2] : 1.0));
            shop_h<2, 0>(acc, A.in, A.g_in, nb, nb, -0.5 * e2 * (wb ? A.bc[2] : 1.0));
        }
        {
            const int st = Xh * A.gh.Y * A.gh.Z;
            const bool wf = (t == A.gh.T - 1), wb = (t == 0);
            const int nf = wf ? h - (A.gh.T - 1) * st : h + st, nb = wb ? h + (A.gh.T - 1) * st : h - st;
            shop_h<3, 1>(acc, A.in, A.g_out, nf, h, 0.5 * e3 * (wf ? A.bc[3] : 1.0));
            shop_h<3, 0>(acc, A.in, A.g_in, nb, nb, -0.5 * e3 * (wb ? A.bc[3] : 1.0));
        }
        const size_t base = (size_t)blk * (3 * 32) + lane;
        cplx *dst = A.fuse.axpy_r ? A.fuse.axpy_r : A.out;
        const double malpha = A.fuse.axpy_r ? -A.red.st->alpha : 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            cplx yk = cmake(A.coef * acc[k].x, A.coef * acc[k].y);
            if (A.xsrc) {
                const cplx xi = ldg128(A.xsrc + base + k * 32);
                yk = cmake(fma(A.coef, acc[k].x, A.xcoef * xi.x), fma(A.coef, acc[k].y, A.xcoef * xi.y));
            }
            if (A.fuse.axpy_r) {           // fused CG residual update: r <- r - alpha * y; y itself is not stored
                const cplx rv = A.fuse.axpy_r[base + k * 32];
                yk = cmake(fma(malpha, yk.x, rv.x), fma(malpha, yk.y, rv.y));
            }
            if (A.fuse.dot_with) {
                const cplx w = ldg128(A.fuse.dot_with + base + k * 32);
                red[0] = fma(w.x, yk.x, red[0]); red[0] = fma(w.y, yk.y, red[0]);
                red[1] = fma(w.x, yk.y, red[1]); red[1] = fma(-w.y, yk.x, red[1]);
            }
            red[2] = fma(yk.x, yk.x, red[2]); red[2] = fma(yk.y, yk.y, red[2]);
            if (A.norm_src) {
                const cplx pv = ldg128(A.norm_src + base + k * 32);
                red[2] = fma(A.norm_coef * pv.x, pv.x, red[2]); red[2] = fma(A.norm_coef * pv.y, pv.y, red[2]);
            }
            dst[base + k * 32] = yk;
        }
    }
    if (A.fuse.dot_with || A.fuse.want_norm) grid_reduce_finish<3>(red, A.red, A.fuse.finish);
}

static int stag_hop(lqcd_ctx *ctx, EoState *e, const lqcd_op *op, int out_parity, cplx *out, const cplx *in, const cplx *xsrc,
                    double coef, double xcoef, const cplx *norm_src, double norm_coef, const DslashFuse *fuse) {
    StagEoArgs A;
    A.out = out; A.in = in; A.xsrc = xsrc; A.norm_src = norm_src; A.coef = coef; A.xcoef = xcoef; A.norm_coef = norm_coef;
    A.g_out = e->gauge[out_parity]; A.g_in = e->gauge[1 - out_parity];
    A.gh = e->gh; A.parity = out_parity;
    for (int i = 0; i < 4; i++) A.bc[i] = op->bc[i];
    A.fuse = fuse ? *fuse : DslashFuse();
    A.red = ctx->red;
    if (A.fuse.shift_src) return lqcd_fail(ctx, LQCD_ERR_ARG, "staggered even-site hop: unsupported fused epilogue");
    const int bs = 32 * e->gh.wpc, grid = (e->gh.nblk + e->gh.wpc - 1) / e->gh.wpc;
    if (bs > 256) return lqcd_fail(ctx, LQCD_ERR_ARG, "LQCD_WPC > 8 is not supported");
    staggered_eo_hop_kernel<<<grid, bs, 0, ctx->stream>>>(A);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}

// the solver's "D" (dagger = 0) and "D^dag" (dagger = 1) while ctx->eo_active is set for a staggered operator: together A_ee
int stag_even_apply(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse) {
    EoState *e = ctx->eo;
    const double m2 = op->mass * op->mass;
    if (!dagger) {
        e->stag_in = x;
        return stag_hop(ctx, e, op, 1, y, x, nullptr, 1.0, 0.0, x, m2, fuse);           // t_o = Dh_oe p_e ; |t|^2 + m^2 |p|^2
    }
    if (!e->stag_in) return lqcd_fail(ctx, LQCD_ERR_STATE, "staggered even-site operator: second half applied before the first");
    const cplx *p = e->stag_in;
    e->stag_in = nullptr;
    return stag_hop(ctx, e, op, 0, y, x, p, -1.0, m2, nullptr, 0.0, fuse);               // q_e = m^2 p_e - Dh_eo t_o
}

// (DdagD) y = b for a source that lives on even sites: CG on A_ee.  The odd sites of b are ignored, the even part of y is the
// initial guess, and y comes back with zeros on the odd sites (where the exact solution of the full system is zero).
extern "C" int lqcd_solve_staggered_even(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *b, double eps, int maxsteps,
                                         int *iters, double *resid_sq) {
    if (!ctx || !op || !y || !b) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (op->kind != LQCD_STAGGERED) return lqcd_fail(ctx, LQCD_ERR_ARG, "the even-site solve is built for the staggered operator");
    if (y->owner != ctx || b->owner != ctx || y->kind != LQCD_STAGGERED || b->kind != LQCD_STAGGERED || y == b) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad fields");
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
    cplx *be = e->f[0], *junk = e->f[1], *xe = e->f[2];
    LQCD_TRY(eo_convert(ctx, e, 1, b->d, be, junk, 3));
    LQCD_TRY(eo_convert(ctx, e, 1, y->d, xe, junk, 3));
    const size_t nhalf = (size_t)e->gh.nblk * 3 * 32;
    e->stag_in = nullptr;
    ctx->eo_active = 1;
    const int rc = solve_impl(ctx, op, xe, be, nhalf, LQCD_SOLVER_CG, LQCD_OP_DDAGD, eps, maxsteps, iters, resid_sq, nullptr);
    ctx->eo_active = 0;
    if (rc != LQCD_OK && rc != LQCD_ERR_NOCONV) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(junk, 0, nhalf * sizeof(cplx), ctx->stream));
    LQCD_TRY(eo_convert(ctx, e, 0, y->d, xe, junk, 3));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

// Scoped routing switch (default off): while set, lqcd_solve(CG, DdagD) calls on a staggered operator -- also the ones issued
// inside lqcd_fermion_force / lqcd_md_trajectory -- go through lqcd_solve_staggered_even.  The caller guarantees even-site sources
// (the host mirrors set it only around the solves of an even-site pseudofermion action).
extern "C" int lqcd_set_staggered_even_solve(lqcd_ctx *ctx, int on) {
    if (!ctx) return lqcd_fail(ctx, LQCD_ERR_ARG, "null ctx");
    ctx->stag_even_solve = on ? 1 : 0;
    return LQCD_OK;
}
