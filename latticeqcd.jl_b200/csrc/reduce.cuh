// reduce.cuh -- deterministic grid-wide reductions with an in-kernel "finish" step.
//
// Every reducing kernel ends with grid_reduce_finish(): warp-shuffle butterfly -> shared memory ->
// one partial per CTA written to global memory -> the LAST CTA to arrive (atomic ticket) sums the
// partials in a fixed order and applies a FinishOp that updates the Krylov scalars in device memory
// (alpha, beta, |r|^2, convergence flag).  No host round trip, no floating-point atomics: the result
// depends only on (gridDim, blockDim), never on scheduling order, so solver iteration counts are
// reproducible run to run.
#pragma once
#include "lqcd_internal.cuh"

__device__ __forceinline__ void apply_finish(int op, const double *tot, SolverState *st, double *hist) {
    switch (op) {
    case FIN_STORE:
        for (int j = 0; j < LQCD_MAX_RED; j++) st->red[j] = tot[j];
        break;
    case FIN_CG_INIT:
        st->rr = tot[0]; st->it = 0;
        if (hist) hist[0] = tot[0];
        if (tot[0] < st->eps) { st->done = 1; st->iters = 0; }
        break;
    case FIN_CG_PQ:
        st->pq = tot[0];
        st->alpha = st->rr / tot[0];
        break;
    case FIN_CG_PQN:
        st->pq = tot[2];
        st->alpha = st->rr / tot[2];
        break;
    case FIN_CG_RR:
    case FIN_CG_RRN: {
        int it = ++st->it;
        double c3 = (op == FIN_CG_RRN) ? tot[2] : tot[0];
        if (hist) hist[it] = c3;
        if (c3 < st->eps) { st->done = 1; st->iters = it; st->rr = c3; }
        else { st->beta = c3 / st->rr; st->rr = c3; }
    } break;
    case FIN_NR_C1:
        st->c1 = tot[2];
        break;
    case FIN_NR_C2:
        st->pq = tot[2];
        st->alpha = st->c1 / tot[2];
        break;
    case FIN_NR_RR: {
        int it = ++st->it;
        st->rr = tot[0];
        if (hist) hist[it] = tot[0];
        if (tot[0] < st->eps) { st->done = 1; st->iters = it; }
    } break;
    case FIN_NR_C3:
        st->beta = tot[2] / st->c1;
        st->c1 = tot[2];
        break;
    case FIN_BI_INIT:
        st->rr = tot[0]; st->it = 0; st->rho_re = tot[0]; st->rho_im = 0.0;
        if (hist) hist[0] = tot[0];
        if (tot[0] < st->eps) { st->done = 1; st->iters = 0; }
        break;
    case FIN_BI_ALPHA: {            // alpha = rho / <r0, v>
        double dr = tot[0], di = tot[1], n = dr * dr + di * di;
        st->alpha_re = (st->rho_re * dr + st->rho_im * di) / n;
        st->alpha_im = (st->rho_im * dr - st->rho_re * di) / n;
    } break;
    case FIN_BI_OMEGA: {            // omega = <t,s> / |t|^2 ; the kernel reduced <s,t> = conj(<t,s>)
        st->omega_re = tot[0] / tot[2];
        st->omega_im = -tot[1] / tot[2];
    } break;
    case FIN_BI_RR: {               // red0 = |r|^2, red1/2 = <r0, r>
        int it = ++st->it;
        st->rr = tot[0];
        if (hist) hist[it] = tot[0];
        if (tot[0] < st->eps) { st->done = 1; st->iters = it; }
        else {
            // beta = (rho_new / rho) * (alpha / omega)
            double ar = tot[1], ai = tot[2];
            double n1 = st->rho_re * st->rho_re + st->rho_im * st->rho_im;
            double q1r = (ar * st->rho_re + ai * st->rho_im) / n1, q1i = (ai * st->rho_re - ar * st->rho_im) / n1;
            double n2 = st->omega_re * st->omega_re + st->omega_im * st->omega_im;
            double q2r = (st->alpha_re * st->omega_re + st->alpha_im * st->omega_im) / n2;
            double q2i = (st->alpha_im * st->omega_re - st->alpha_re * st->omega_im) / n2;
            st->bre = q1r * q2r - q1i * q2i;
            st->bim = q1r * q2i + q1i * q2r;
            st->rho_re = ar; st->rho_im = ai;
        }
    } break;
    case FIN_MS_PQ: {               // multi-shift CG (Jegerlehner): base alpha, then zeta / alpha_j
        double alpha = st->rr / tot[0];
        st->pq = tot[0];
        st->alpha = alpha;
        st->alpha_s[0] = alpha;
        for (int j = 1; j < st->nshift; j++) {
            double ds = st->shift[j] - st->shift[0];
            double z = st->zeta[j], zo = st->zeta_old[j];
            // heavily shifted systems converge early: their zeta underflows -> freeze (same rule as the oracle)
            if (fabs(z) < 1e-140) { st->zeta[j] = st->zeta_old[j] = 0.0; st->alpha_s[j] = 0.0; continue; }
            double znew = z * zo * st->alpha_old /
                          (alpha * st->beta_old * (zo - z) + zo * st->alpha_old * (1.0 + ds * alpha));
            st->alpha_s[j] = alpha * znew / z;
            st->zeta_old[j] = z; st->zeta[j] = znew;
        }
    } break;
    case FIN_MS_RR: {
        int it = ++st->it;
        double c3 = tot[0];
        if (hist) hist[it] = c3;
        if (c3 < st->eps) { st->done = 1; st->iters = it; st->rr = c3; }
        else {
            double beta = c3 / st->rr;
            st->beta = beta; st->beta_s[0] = beta;
            for (int j = 1; j < st->nshift; j++) {
                if (st->zeta[j] == 0.0) { st->beta_s[j] = 0.0; continue; }      // frozen shift: p_j <- 0
                double ratio = st->zeta[j] / st->zeta_old[j];
                st->beta_s[j] = beta * ratio * ratio;
            }
            st->alpha_old = st->alpha; st->beta_old = beta; st->rr = c3;
        }
    } break;
    default: break;
    }
}

// NR values per thread.  Must be called by ALL threads of ALL CTAs of the grid (no early return before it).
// Split reductions (multi-GPU Dslash: interior kernel + exterior kernel contribute to ONE sum):
//   part_offset  index of this grid's first partial in R.partials
//   part_total   number of partials the finishing CTA sums (own grid + earlier kernels' partials)
//   do_finish    0: only deposit the CTA partial (a later kernel finishes)
template <int NR>
__device__ __forceinline__ void grid_reduce_finish(double (&v)[NR], const Reduce &R, int finish,
                                                   unsigned int part_offset = 0, unsigned int part_total = 0, int do_finish = 1,
                                                   unsigned int cta_index = 0xffffffffu, unsigned int n_ctas = 0) {
    // cta_index / n_ctas: the CTAs that take part (default: the whole grid); self-packing Dslash kernels exclude
    // their leading pack CTAs.
    if (n_ctas == 0) { n_ctas = gridDim.x; cta_index = blockIdx.x; }
    if (part_total == 0) part_total = n_ctas;
    __shared__ double sm[NR][32];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int j = 0; j < NR; j++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
        if (lane == 0) sm[j][warp] = v[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NR; j++) {
            double s = 0.0;
            for (int w = 0; w < nwarp; w++) s += sm[j][w];
            R.partials[(size_t)(part_offset + cta_index) * LQCD_MAX_RED + j] = s;
        }
        is_last = 0;
        if (do_finish) {
            __threadfence();
            unsigned int t = atomicInc(R.ticket, n_ctas - 1);
            is_last = (t == n_ctas - 1);
        }
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double acc[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) acc[j] = 0.0;
    for (unsigned int i = threadIdx.x; i < part_total; i += blockDim.x) {
#pragma unroll
        for (int j = 0; j < NR; j++) acc[j] += __ldcg(&R.partials[(size_t)i * LQCD_MAX_RED + j]);
    }
#pragma unroll
    for (int j = 0; j < NR; j++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NR; j++) sm[j][warp] = acc[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot[LQCD_MAX_RED];
#pragma unroll
        for (int j = 0; j < LQCD_MAX_RED; j++) tot[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NR; j++) {
            double s = 0.0;
            for (int w = 0; w < nwarp; w++) s += sm[j][w];
            tot[j] = s;
        }
        if (R.cr.nranks > 1) {      // cross-GPU all-reduce over peer memory, deterministic rank order
            const CommRed &c = R.cr;
            const unsigned long long seq = ++(*c.seq);
            const int slot = (int)(seq % LQCD_RED_SLOTS);
            for (int r = 0; r < c.nranks; r++) {
                double *dst = c.vals[r] + ((size_t)slot * c.nranks + c.rank) * LQCD_MAX_RED;
#pragma unroll
                for (int j = 0; j < NR; j++) dst[j] = tot[j];
            }
            __threadfence_system();
            for (int r = 0; r < c.nranks; r++) st_release_sys(c.flags[r] + (size_t)slot * c.nranks + c.rank, seq);
            const long long t0 = clock64();
            bool ok = true;
            for (int r = 0; r < c.nranks && ok; r++) {
                while (ld_acquire_sys(c.flags[c.rank] + (size_t)slot * c.nranks + r) < seq) {
                    if (clock64() - t0 > c.timeout_cycles) { ok = false; break; }
                }
            }
            if (!ok) {       // a lost peer: report LQCD_ERR_COMM (failed = 2) and stop here -- the partial sums must not reach the finish op
                *c.err = 2000000 + (int)(seq % 1000000); R.st->done = 1; R.st->failed = 2; R.st->iters = R.st->it;
                return;
            }
#pragma unroll
            for (int j = 0; j < NR; j++) {
                double s = 0.0;
                for (int r = 0; r < c.nranks; r++)
                    s += __ldcg(c.vals[c.rank] + ((size_t)slot * c.nranks + r) * LQCD_MAX_RED + j);
                tot[j] = s;
            }
        }
        apply_finish(finish, tot, R.st, R.hist);
        if (finish != FIN_STORE && !(fabs(tot[0]) <= 1.79e308 && fabs(tot[2]) <= 1.79e308)) { R.st->done = 1; R.st->failed = 1; R.st->iters = R.st->it; }
    }
}
