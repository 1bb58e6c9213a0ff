/*
 * lqcd_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  *** PARITY UNPINNED ***
 * See lqcd_oracle.h for scope, layouts and the pinning statement.
 *
 * Each routine cites the reference call site it serves (paths relative to /root/reference) and the
 * SURVEY.md appendix that restates the upstream (LatticeDiracOperators.jl 0.6.x) algorithm it follows.
 */
#include "lqcd_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}

/* ---- gamma matrices: upstream WilsonFermion constructor tables (recalled) ---- */
void orc_set_gamma(orc_op *op, double r) {
    zc g[4][4][4];
    memset(g, 0, sizeof g);
    /* 0-based [mu][row][col]; the 1-based table in SURVEY.md 8c */
    g[0][0][3] = -I; g[0][1][2] = -I; g[0][2][1] =  I; g[0][3][0] =  I;
    g[1][0][3] = -1; g[1][1][2] =  1; g[1][2][1] =  1; g[1][3][0] = -1;
    g[2][0][2] = -I; g[2][1][3] =  I; g[2][2][0] =  I; g[2][3][1] = -I;
    g[3][0][2] = -1; g[3][1][3] = -1; g[3][2][0] = -1; g[3][3][1] = -1;
    for (int mu = 0; mu < 4; mu++)
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++) {
                zc id = (a == b) ? r : 0.0;
                op->rplusg[mu][a][b]  = id + g[mu][a][b];
                op->rminusg[mu][a][b] = id - g[mu][a][b];
            }
    op->r = r;
}

/* ---- geometry ---- */
typedef struct { int64_t V; int d[4]; int64_t stride[4]; } geom;
static geom mkgeom(const int dims[4]) {
    geom g; g.V = 1;
    for (int i = 0; i < 4; i++) { g.d[i] = dims[i]; g.stride[i] = g.V; g.V *= dims[i]; }
    return g;
}
static inline void site_coords(const geom *g, int64_t s, int c[4]) {
    for (int i = 0; i < 4; i++) { c[i] = (int)(s % g->d[i]); s /= g->d[i]; }
}
/* neighbour in +/-mu with wrap flag */
static inline int64_t nbr(const geom *g, int64_t s, const int c[4], int mu, int sign, int *wrapped) {
    if (sign > 0) {
        if (c[mu] == g->d[mu] - 1) { *wrapped = 1; return s - (int64_t)(g->d[mu] - 1) * g->stride[mu]; }
        *wrapped = 0; return s + g->stride[mu];
    } else {
        if (c[mu] == 0) { *wrapped = 1; return s + (int64_t)(g->d[mu] - 1) * g->stride[mu]; }
        *wrapped = 0; return s - g->stride[mu];
    }
}

/* ---- Wilson hopping, LinearAlgebra.mul!(y, D, x) -> upstream Wx!
 *      served call sites: AbstractMD.jl:129 (inside calc_UdSfdU!), standardHMC.jl:69-71,
 *      measurements/unusedfiles/measure_Pion_correlator.jl:379,399.
 *  Per direction nu, in the reference's order of operations:
 *      temp1 = (r - gamma_nu) [U_nu(n) x(n+nu)],  temp2 = (r + gamma_nu) [U_nu^dag(n-nu) x(n-nu)],
 *      acc  += kappa*temp1 + kappa*temp2;   finally y = x - acc.
 *  dagger swaps (r-gamma) <-> (r+gamma)  (upstream Wdagx!).  Boundary phase bc[nu] multiplies the shifted
 *  spinor when the shift wraps (upstream applies it when filling the fermion "wing").                    */
/* parity < 0: all sites, y = [A] x - kappa H x.  parity = 0/1: only sites of that parity are computed and the result is
 * the bare hopping sum y = H x there (kappa not applied), zero on the other parity (even-odd building block). */
static void wilson_apply_p(const orc_op *op, int dagger, zc *y, const zc *const u[4], const zc *x, int parity);
static void wilson_apply(const orc_op *op, int dagger, zc *y, const zc *const u[4], const zc *x) {
    wilson_apply_p(op, dagger, y, u, x, -1);
}
void orc_wilson_hop_parity(const orc_op *op, int dagger, int parity, zc *y, const zc *const u[4], const zc *x) {
    wilson_apply_p(op, dagger, y, u, x, parity & 1);
}
static void wilson_apply_p(const orc_op *op, int dagger, zc *y, const zc *const u[4], const zc *x, int parity) {
    geom g = mkgeom(op->dims);
    if (op->csw != 0.0 && !op->clov) abort();      /* orc_clover_build first */
    const int64_t V = g.V;
    const zc (*Gf)[4][4] = dagger ? op->rplusg : op->rminusg;   /* multiplies the forward hop */
    const zc (*Gb)[4][4] = dagger ? op->rminusg : op->rplusg;   /* multiplies the backward hop */
    const double kappa = parity < 0 ? op->kappa : 1.0;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < V; s++) {
        int c[4]; site_coords(&g, s, c);
        zc acc[4][3];
        memset(acc, 0, sizeof acc);
        if (parity >= 0) {
            if (((c[0] + c[1] + c[2] + c[3]) & 1) != parity) {
                for (int al = 0; al < 4; al++) for (int a = 0; a < 3; a++) y[a + 3 * (s + V * al)] = 0;
                continue;
            }
        }
        for (int nu = 0; nu < 4; nu++) {
            int wf, wb;
            int64_t sf = nbr(&g, s, c, nu, +1, &wf);
            int64_t sb = nbr(&g, s, c, nu, -1, &wb);
            double pf = wf ? op->bc[nu] : 1.0, pb = wb ? op->bc[nu] : 1.0;
            const zc *Uf = u[nu] + 9 * s;     /* U_nu(n)      [a + 3 b] */
            const zc *Ub = u[nu] + 9 * sb;    /* U_nu(n - nu)           */
            zc h[4][3], gq[4][3];
            for (int al = 0; al < 4; al++) {
                zc xf[3], xb[3];
                for (int b = 0; b < 3; b++) {
                    xf[b] = pf * x[b + 3 * (sf + V * al)];
                    xb[b] = pb * x[b + 3 * (sb + V * al)];
                }
                for (int a = 0; a < 3; a++) {
                    zc sfw = 0, sbw = 0;
                    for (int b = 0; b < 3; b++) {
                        sfw += Uf[a + 3 * b] * xf[b];             /* U x          */
                        sbw += conj(Ub[b + 3 * a]) * xb[b];       /* U^dag x      */
                    }
                    h[al][a] = sfw; gq[al][a] = sbw;
                }
            }
            for (int al = 0; al < 4; al++)
                for (int a = 0; a < 3; a++) {
                    zc t1 = 0, t2 = 0;
                    for (int be = 0; be < 4; be++) {
                        t1 += Gf[nu][al][be] * h[be][a];
                        t2 += Gb[nu][al][be] * gq[be][a];
                    }
                    acc[al][a] += kappa * t1 + kappa * t2;
                }
        }
        if (parity >= 0) {
            for (int al = 0; al < 4; al++) for (int a = 0; a < 3; a++) y[a + 3 * (s + V * al)] = acc[al][a];
        } else
        if (op->csw != 0.0) {       /* Wilson-clover: y = A(n) x(n) - acc, A block diagonal in chirality (spins 01 | 23) */
            const zc *A = op->clov + 72 * s;
            for (int blk = 0; blk < 2; blk++)
                for (int i = 0; i < 6; i++) {
                    zc sum = 0;
                    for (int j = 0; j < 6; j++)
                        sum += A[36 * blk + i + 6 * j] * x[(j % 3) + 3 * (s + V * (2 * blk + j / 3))];
                    y[(i % 3) + 3 * (s + V * (2 * blk + i / 3))] = sum - acc[2 * blk + i / 3][i % 3];
                }
        } else
        for (int al = 0; al < 4; al++)
            for (int a = 0; a < 3; a++)
                y[a + 3 * (s + V * al)] = x[a + 3 * (s + V * al)] - acc[al][a];
    }
}

/* ---- staggered, mul!(y, D, x) -> upstream Dx! + mass
 *      y = m x + sum_nu (1/2) eta_nu(n) [U_nu(n) x(n+nu) - U_nu^dag(n-nu) x(n-nu)],
 *      eta_1 = 1, eta_2 = (-1)^x, eta_3 = (-1)^(x+y), eta_4 = (-1)^(x+y+z);  D^dag = m - hop.          */
static void staggered_apply(const orc_op *op, int dagger, zc *y, const zc *const u[4], const zc *x) {
    geom g = mkgeom(op->dims);
    const int64_t V = g.V;
    const double sgn = dagger ? -1.0 : 1.0;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < V; s++) {
        int c[4]; site_coords(&g, s, c);
        zc acc[3] = {0, 0, 0};
        int esum = 0;
        for (int nu = 0; nu < 4; nu++) {
            double eta = (esum & 1) ? -1.0 : 1.0;
            esum += c[nu];
            int wf, wb;
            int64_t sf = nbr(&g, s, c, nu, +1, &wf);
            int64_t sb = nbr(&g, s, c, nu, -1, &wb);
            double pf = wf ? op->bc[nu] : 1.0, pb = wb ? op->bc[nu] : 1.0;
            const zc *Uf = u[nu] + 9 * s, *Ub = u[nu] + 9 * sb;
            for (int a = 0; a < 3; a++) {
                zc t1 = 0, t2 = 0;
                for (int b = 0; b < 3; b++) {
                    t1 += Uf[a + 3 * b] * (pf * x[b + 3 * sf]);
                    t2 += conj(Ub[b + 3 * a]) * (pb * x[b + 3 * sb]);
                }
                acc[a] += eta * (0.5 * t1 - 0.5 * t2);
            }
        }
        for (int a = 0; a < 3; a++) y[a + 3 * s] = op->mass * x[a + 3 * s] + sgn * acc[a];
    }
}

static int64_t field_len(const orc_op *op, int kind) {
    int64_t V = (int64_t)op->dims[0] * op->dims[1] * op->dims[2] * op->dims[3];
    return kind == ORC_STAGGERED ? 3 * V : 12 * V;
}

void orc_apply(const orc_op *op, int kind, int mode, zc *y, const zc *const u[4], const zc *x, zc *scratch) {
    if (mode == ORC_DDAGD) {   /* upstream DdagD: y = D^dag (D x) through one scratch field */
        zc *t = scratch ? scratch : (zc *)malloc(sizeof(zc) * field_len(op, kind));
        orc_apply(op, kind, ORC_D, t, u, x, NULL);
        orc_apply(op, kind, ORC_DDAG, y, u, t, NULL);
        if (!scratch) free(t);
        return;
    }
    if (kind == ORC_WILSON) wilson_apply(op, mode == ORC_DDAG, y, u, x);
    else if (kind == ORC_WILSON_EO) {   /* y_e = x_e - kappa^2 H_eo (H_oe x_e)  (dagger: H^dag in both hops) */
        const int64_t n = field_len(op, kind);
        zc *t = malloc(sizeof(zc) * n);
        wilson_apply_p(op, mode == ORC_DDAG, t, u, x, 1);
        wilson_apply_p(op, mode == ORC_DDAG, y, u, t, 0);
        const double k2 = op->kappa * op->kappa;
        geom g = mkgeom(op->dims);
#pragma omp parallel for schedule(static)
        for (int64_t s = 0; s < g.V; s++) {
            int c[4]; site_coords(&g, s, c);
            const int even = ((c[0] + c[1] + c[2] + c[3]) & 1) == 0;
            for (int al = 0; al < 4; al++) for (int a = 0; a < 3; a++) {
                int64_t i = a + 3 * (s + g.V * al);
                y[i] = even ? x[i] - k2 * y[i] : 0;
            }
        }
        free(t);
    }
    else                    staggered_apply(op, mode == ORC_DDAG, y, u, x);
}

/* ---- BLAS-1 (upstream add!, dot; standardHMC.jl:54 dot(xi,xi)) ---- */
zc orc_dot(const zc *a, const zc *b, int64_t n) {
    double re = 0, im = 0;
#pragma omp parallel for reduction(+ : re, im) schedule(static)
    for (int64_t i = 0; i < n; i++) { zc t = conj(a[i]) * b[i]; re += creal(t); im += cimag(t); }
    return re + im * I;
}
static void axpy(zc *y, zc a, const zc *x, int64_t n) {        /* y += a x   (add!(y, a, x)) */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) y[i] += a * x[i];
}
static void xpby(zc *y, const zc *x, zc b, int64_t n) {        /* y = b y + x (add!(b, y, 1, x)) */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) y[i] = b * y[i] + x[i];
}
static void copyv(zc *y, const zc *x, int64_t n) { memcpy(y, x, sizeof(zc) * n); }

/* ---- CG on A = D^dag D: solve_DinvX!(y, DdagD, x)
 *      call sites: calc_UdSfdU! (AbstractMD.jl:129), evaluate_FermiAction (standardHMC.jl:69-71).      */
int orc_cg(const orc_op *op, int kind, zc *x, const zc *const u[4], const zc *b,
           double eps, int maxsteps, double *resid_sq, double *hist) {
    const int64_t n = field_len(op, kind);
    zc *res = malloc(sizeof(zc) * n), *q = malloc(sizeof(zc) * n), *p = malloc(sizeof(zc) * n),
       *t = malloc(sizeof(zc) * n);
    int ret = -1;
    orc_apply(op, kind, ORC_DDAGD, q, u, x, t);
    copyv(res, b, n); axpy(res, -1.0, q, n);
    copyv(p, res, n);
    double rnorm = creal(orc_dot(res, res, n));
    if (hist) hist[0] = rnorm;
    if (rnorm < eps) { ret = 0; goto done; }
    for (int i = 1; i <= maxsteps; i++) {
        orc_apply(op, kind, ORC_DDAGD, q, u, p, t);
        double c1 = creal(orc_dot(p, q, n));
        double alpha = rnorm / c1;
        axpy(x, alpha, p, n);
        axpy(res, -alpha, q, n);
        double c3 = creal(orc_dot(res, res, n));
        if (hist) hist[i] = c3;
        if (c3 < eps) { rnorm = c3; ret = i; goto done; }
        double beta = c3 / rnorm;
        xpby(p, res, beta, n);
        rnorm = c3;
    }
done:
    if (resid_sq) *resid_sq = rnorm;
    free(res); free(q); free(p); free(t);
    return ret;
}

/* ---- upstream "bicg" = CGNR on A = D: solve_DinvX!(y, D, x)
 *      call sites: measure_Pion_correlator.jl:399, measure_chiral_condensate.jl:182 (via QCDMeasurements). */
int orc_cgnr(const orc_op *op, int kind, zc *x, const zc *const u[4], const zc *b,
             double eps, int maxsteps, double *resid_sq, double *hist) {
    const int64_t n = field_len(op, kind);
    zc *res = malloc(sizeof(zc) * n), *q = malloc(sizeof(zc) * n), *p = malloc(sizeof(zc) * n);
    int ret = -1;
    orc_apply(op, kind, ORC_D, q, u, x, NULL);
    copyv(res, b, n); axpy(res, -1.0, q, n);
    double rnorm = creal(orc_dot(res, res, n));
    if (hist) hist[0] = rnorm;
    if (rnorm < eps) { ret = 0; goto done; }
    orc_apply(op, kind, ORC_DDAG, q, u, res, NULL);
    copyv(p, q, n);
    double c1 = creal(orc_dot(q, q, n));
    for (int i = 1; i <= maxsteps; i++) {
        orc_apply(op, kind, ORC_D, q, u, p, NULL);
        double c2 = creal(orc_dot(q, q, n));
        double alpha = c1 / c2;
        axpy(res, -alpha, q, n);
        axpy(x, alpha, p, n);
        rnorm = creal(orc_dot(res, res, n));
        if (hist) hist[i] = rnorm;
        if (rnorm < eps) { ret = i; goto done; }
        orc_apply(op, kind, ORC_DDAG, q, u, res, NULL);
        double c3 = creal(orc_dot(q, q, n));
        double beta = c3 / c1;
        c1 = c3;
        xpby(p, q, beta, n);
    }
done:
    if (resid_sq) *resid_sq = rnorm;
    free(res); free(q); free(p);
    return ret;
}

/* ---- BiCGStab on A = D (params["method_CG"]="bicgstab", SURVEY.md App. C.4 last line).
 *      Textbook van der Vorst recurrences with shadow residual r0~ = r0; same stopping rule.           */
int orc_bicgstab(const orc_op *op, int kind, zc *x, const zc *const u[4], const zc *b,
                 double eps, int maxsteps, double *resid_sq, double *hist) {
    const int64_t n = field_len(op, kind);
    zc *r = malloc(sizeof(zc) * n), *r0 = malloc(sizeof(zc) * n), *p = malloc(sizeof(zc) * n),
       *v = malloc(sizeof(zc) * n), *s = malloc(sizeof(zc) * n), *t = malloc(sizeof(zc) * n);
    int ret = -1;
    orc_apply(op, kind, ORC_D, v, u, x, NULL);
    copyv(r, b, n); axpy(r, -1.0, v, n);
    copyv(r0, r, n); copyv(p, r, n);
    double rnorm = creal(orc_dot(r, r, n));
    if (hist) hist[0] = rnorm;
    if (rnorm < eps) { ret = 0; goto done; }
    zc rho = orc_dot(r0, r, n);
    for (int i = 1; i <= maxsteps; i++) {
        orc_apply(op, kind, ORC_D, v, u, p, NULL);
        zc alpha = rho / orc_dot(r0, v, n);
        copyv(s, r, n); axpy(s, -alpha, v, n);
        orc_apply(op, kind, ORC_D, t, u, s, NULL);
        zc omega = orc_dot(t, s, n) / creal(orc_dot(t, t, n));
        axpy(x, alpha, p, n); axpy(x, omega, s, n);
        copyv(r, s, n); axpy(r, -omega, t, n);
        rnorm = creal(orc_dot(r, r, n));
        if (hist) hist[i] = rnorm;
        if (rnorm < eps) { ret = i; goto done; }
        zc rho_new = orc_dot(r0, r, n);
        zc beta = (rho_new / rho) * (alpha / omega);
        rho = rho_new;
        /* p = r + beta (p - omega v) */
        axpy(p, -omega, v, n);
        xpby(p, r, beta, n);
    }
done:
    if (resid_sq) *resid_sq = rnorm;
    free(r); free(r0); free(p); free(v); free(s); free(t);
    return ret;
}

/* ---- multi-shift CG (upstream shiftedcg): (D^dag D + sigma_j) x_j = b.
 *      Used by RHMC (test/test_Nf2.toml, universe.jl:106-110).  Zero initial guess; zeta recurrences
 *      (Jegerlehner hep-lat/9612014); convergence tested on the residual of shifts[0].                  */
int orc_mscg(const orc_op *op, int kind, zc *const xs[], const zc *const u[4], const zc *b,
             const double *shifts, int nshift, double eps, int maxsteps, double *resid_sq) {
    const int64_t n = field_len(op, kind);
    zc *r = malloc(sizeof(zc) * n), *q = malloc(sizeof(zc) * n), *t = malloc(sizeof(zc) * n);
    zc **ps = malloc(sizeof(zc *) * nshift);
    double *zeta = malloc(sizeof(double) * nshift), *zeta_old = malloc(sizeof(double) * nshift),
           *beta_s = malloc(sizeof(double) * nshift);
    for (int j = 0; j < nshift; j++) {
        ps[j] = malloc(sizeof(zc) * n);
        copyv(ps[j], b, n);
        memset(xs[j], 0, sizeof(zc) * n);
        zeta[j] = zeta_old[j] = 1.0;
    }
    copyv(r, b, n);
    int ret = -1;
    double rr = creal(orc_dot(r, r, n));
    double alpha_old = 1.0, beta_old = 0.0;   /* CG scalars of the base system in "x += alpha p" form */
    if (rr < eps) { ret = 0; goto done; }
    const double s0 = shifts[0];
    for (int i = 1; i <= maxsteps; i++) {
        /* base system A0 = D^dag D + s0 */
        orc_apply(op, kind, ORC_DDAGD, q, u, ps[0], t);
        axpy(q, s0, ps[0], n);
        double pq = creal(orc_dot(ps[0], q, n));
        double alpha = rr / pq;
        /* shifted coefficients (relative shift ds = sigma_j - s0) */
        for (int j = 1; j < nshift; j++) {
            double ds = shifts[j] - s0;
            /* a heavily shifted system converges long before the base one: its zeta underflows and the recurrence
             * would produce 0/0.  Freeze it (its residual zeta_j^2 |r|^2 is far below eps by then). */
            if (fabs(zeta[j]) < 1e-140) { zeta_old[j] = zeta[j] = 0.0; continue; }
            double znew = zeta[j] * zeta_old[j] * alpha_old /
                          (alpha * beta_old * (zeta_old[j] - zeta[j]) + zeta_old[j] * alpha_old * (1.0 + ds * alpha));
            double alpha_j = alpha * znew / zeta[j];
            axpy(xs[j], alpha_j, ps[j], n);
            zeta_old[j] = zeta[j]; zeta[j] = znew;
            beta_s[j] = alpha_j;   /* stash alpha_j for the beta_j update */
        }
        axpy(xs[0], alpha, ps[0], n);
        axpy(r, -alpha, q, n);
        double rr_new = creal(orc_dot(r, r, n));
        if (rr_new < eps) { rr = rr_new; ret = i; goto done; }
        double beta = rr_new / rr;
        xpby(ps[0], r, beta, n);
        for (int j = 1; j < nshift; j++) {
            /* beta_j = beta * (zeta_new/zeta_old_iter)^2 ; p_j = zeta_new r + beta_j p_j */
            if (zeta[j] == 0.0) continue;            /* frozen shift */
            double ratio = zeta[j] / zeta_old[j];
            double beta_j = beta * ratio * ratio;
            zc *pj = ps[j];
            const double zj = zeta[j];
#pragma omp parallel for schedule(static)
            for (int64_t k = 0; k < n; k++) pj[k] = beta_j * pj[k] + zj * r[k];
        }
        alpha_old = alpha; beta_old = beta; rr = rr_new;
    }
done:
    if (resid_sq) *resid_sq = rr;
    for (int j = 0; j < nshift; j++) free(ps[j]);
    free(ps); free(zeta); free(zeta_old); free(beta_s); free(r); free(q); free(t);
    return ret;
}

/* ---- plaquette (pins the fixture loader + link layout; SURVEY.md section 4 values) ---- */
static void mm(zc *c, const zc *a, const zc *b) {          /* c = a b, [row + 3 col] */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        zc s = 0; for (int k = 0; k < 3; k++) s += a[i + 3 * k] * b[k + 3 * j];
        c[i + 3 * j] = s;
    }
}
static void mmd(zc *c, const zc *a, const zc *b) {         /* c = a b^dag */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        zc s = 0; for (int k = 0; k < 3; k++) s += a[i + 3 * k] * conj(b[j + 3 * k]);
        c[i + 3 * j] = s;
    }
}
double orc_plaquette(const int dims[4], const zc *const u[4]) {
    geom g = mkgeom(dims);
    double sum = 0;
#pragma omp parallel for reduction(+ : sum) schedule(static)
    for (int64_t s = 0; s < g.V; s++) {
        int c[4]; site_coords(&g, s, c);
        for (int mu = 0; mu < 4; mu++)
            for (int nu = mu + 1; nu < 4; nu++) {
                int w;
                int64_t smu = nbr(&g, s, c, mu, +1, &w), snu = nbr(&g, s, c, nu, +1, &w);
                zc a[9], b[9], d[9];
                mm(a, u[mu] + 9 * s, u[nu] + 9 * smu);      /* U_mu(n) U_nu(n+mu) */
                mmd(b, a, u[mu] + 9 * snu);                  /* ... U_mu(n+nu)^dag */
                mmd(d, b, u[nu] + 9 * s);                    /* ... U_nu(n)^dag    */
                sum += creal(d[0] + d[4] + d[8]);
            }
    }
    return sum / (6.0 * 3.0 * (double)g.V);
}

/* ---- Wilson pseudofermion force, calc_UdSfdU! (AbstractMD.jl:129) ----
 *  UdSfdU_mu(n)[a][b] = -kappa * sum_spin [ ((r-gamma_mu) U_mu(n) X(n+mu))_a  conj(Y(n)_b) ]
 *                       +kappa * sum_spin [ X(n)_a  conj( ((r+gamma_mu) U_mu(n) Y(n+mu)) )_b ... ]
 *  written so that dS_f/d eps = -2 Re tr[A * UdSfdU_mu(n)] for U_mu(n) -> exp(eps A) U_mu(n)
 *  (S_f = phi^dag (D^dag D)^-1 phi).  Derivation in DESIGN.md; verified by finite differences in tests. */
void orc_wilson_force(const orc_op *op, zc *const out[4], const zc *const u[4], const zc *X, const zc *Y) {
    geom g = mkgeom(op->dims);
    const int64_t V = g.V;
    const double kappa = op->kappa;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < V; s++) {
        int c[4]; site_coords(&g, s, c);
        for (int mu = 0; mu < 4; mu++) {
            int wf; int64_t sf = nbr(&g, s, c, mu, +1, &wf);
            double pf = wf ? op->bc[mu] : 1.0;
            const zc *U = u[mu] + 9 * s;
            /* hX = U X(n+mu), hY = U Y(n+mu) (phase included), per spin */
            zc hX[4][3], hY[4][3];
            for (int al = 0; al < 4; al++)
                for (int a = 0; a < 3; a++) {
                    zc sx = 0, sy = 0;
                    for (int b = 0; b < 3; b++) {
                        sx += U[a + 3 * b] * (pf * X[b + 3 * (sf + V * al)]);
                        sy += U[a + 3 * b] * (pf * Y[b + 3 * (sf + V * al)]);
                    }
                    hX[al][a] = sx; hY[al][a] = sy;
                }
            /* delta D = -kappa [ (r-g) dU X(n+mu) at n  +  (r+g) dU^dag X(n) at n+mu ]
             * dS = -2 Re( Y^dag dD X ).  With dU = A U :
             *   term1: Y(n)^dag (r-g) A U X(n+mu)         -> tr A * [ (r-g) hX ] Y(n)^dag
             *   term2: Y(n+mu)^dag (r+g) U^dag A^dag X(n) = -[ (U (r+g) Y(n+mu))^dag A X(n) ] -> tr A * X(n) [(r+g) hY]^dag (sign -)
             * => dS = -2 Re tr A * { -kappa (r-g)hX Y(n)^dag + kappa X(n) ((r+g) hY)^dag }                */
            zc pX[4][3], pY[4][3];
            for (int al = 0; al < 4; al++)
                for (int a = 0; a < 3; a++) {
                    zc t1 = 0, t2 = 0;
                    for (int be = 0; be < 4; be++) {
                        t1 += op->rminusg[mu][al][be] * hX[be][a];
                        t2 += op->rplusg[mu][al][be] * hY[be][a];
                    }
                    pX[al][a] = t1; pY[al][a] = t2;
                }
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) {
                    zc m = 0;
                    for (int al = 0; al < 4; al++) {
                        m += -kappa * pX[al][a] * conj(Y[b + 3 * (s + V * al)]);
                        m +=  kappa * X[a + 3 * (s + V * al)] * conj(pY[al][b]);
                    }
                    out[mu][a + 3 * (b + 3 * s)] = m;
                }
        }
    }
}

/* staggered analogue: delta D = (1/2) eta [ dU X(n+mu) at n - dU^dag X(n) at n+mu ] */
void orc_staggered_force(const orc_op *op, zc *const out[4], const zc *const u[4], const zc *X, const zc *Y) {
    geom g = mkgeom(op->dims);
    const int64_t V = g.V;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < V; s++) {
        int c[4]; site_coords(&g, s, c);
        int esum = 0;
        for (int mu = 0; mu < 4; mu++) {
            double eta = (esum & 1) ? -1.0 : 1.0;
            esum += c[mu];
            int wf; int64_t sf = nbr(&g, s, c, mu, +1, &wf);
            double pf = wf ? op->bc[mu] : 1.0;
            const zc *U = u[mu] + 9 * s;
            zc hX[3], hY[3];
            for (int a = 0; a < 3; a++) {
                zc sx = 0, sy = 0;
                for (int b = 0; b < 3; b++) {
                    sx += U[a + 3 * b] * (pf * X[b + 3 * sf]);
                    sy += U[a + 3 * b] * (pf * Y[b + 3 * sf]);
                }
                hX[a] = sx; hY[a] = sy;
            }
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++)
                    out[mu][a + 3 * (b + 3 * s)] =
                        0.5 * eta * (hX[a] * conj(Y[b + 3 * s]) + X[a + 3 * s] * conj(hY[b]));
        }
    }
}

/* ---- clover term (Wilson-clover; NEW capability, not reachable from run_LQCD at this commit: universe.jl:106-131,
 *      parameter parsed at parameter_structs.jl:125).  In-tree evidence used: the four-leaf clover of the (dead)
 *      topological-charge code, src/measurements/unusedfiles/measure_topological_charge.jl:299-309 (leaf paths),
 *      :177-200 (traceless anti-Hermitian part, /numofloops=4).  Normalisation of the term itself: see orc_op.csw. */
static void mulx(zc *c, const zc *a, int adja, const zc *b, int adjb) {      /* c = op(a) op(b), [row + 3 col] */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        zc s = 0;
        for (int k = 0; k < 3; k++) {
            zc av = adja ? conj(a[k + 3 * i]) : a[i + 3 * k];
            zc bv = adjb ? conj(b[j + 3 * k]) : b[k + 3 * j];
            s += av * bv;
        }
        c[i + 3 * j] = s;
    }
}
static int64_t hopsite(const geom *g, int64_t s, int c[4], int mu, int sign) {   /* moves c[] too */
    int w; int64_t ns = nbr(g, s, c, mu, sign, &w);
    c[mu] = (c[mu] + sign + g->d[mu]) % g->d[mu];
    return ns;
}
/* ordered product along a closed path starting at n: steps (dir, +1/-1) */
static void path_product(const geom *g, const zc *const u[4], int64_t s, const int c0[4], const int dir[4], const int sgn[4], zc out[9]) {
    int c[4] = {c0[0], c0[1], c0[2], c0[3]};
    zc acc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tmp[9];
    for (int k = 0; k < 4; k++) {
        if (sgn[k] > 0) { mulx(tmp, acc, 0, u[dir[k]] + 9 * s, 0); s = hopsite(g, s, c, dir[k], +1); }
        else            { s = hopsite(g, s, c, dir[k], -1); mulx(tmp, acc, 0, u[dir[k]] + 9 * s, 1); }
        memcpy(acc, tmp, sizeof acc);
    }
    memcpy(out, acc, sizeof acc);
}
void orc_clover_build(const orc_op *op, zc *clov, zc *fmunu, const zc *const u[4]) {
    geom g = mkgeom(op->dims);
    const double coef = op->kappa * op->csw;
    /* sigma_mu_nu = (i/2)[g_mu, g_nu] from the operator's own gamma tables */
    zc gam[4][4][4], sig[6][4][4];
    for (int mu = 0; mu < 4; mu++) for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++)
        gam[mu][a][b] = 0.5 * (op->rplusg[mu][a][b] - op->rminusg[mu][a][b]);
    int pl = 0;
    for (int mu = 0; mu < 4; mu++) for (int nu = mu + 1; nu < 4; nu++, pl++)
        for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) {
            zc s = 0;
            for (int k = 0; k < 4; k++) s += gam[mu][a][k] * gam[nu][k][b] - gam[nu][a][k] * gam[mu][k][b];
            sig[pl][a][b] = 0.5 * I * s;
        }
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < g.V; s++) {
        int c[4]; site_coords(&g, s, c);
        zc F[6][9];
        int p = 0;
        for (int mu = 0; mu < 4; mu++) for (int nu = mu + 1; nu < 4; nu++, p++) {
            const int d1[4] = {mu, nu, mu, nu}, s1[4] = {+1, +1, -1, -1};
            const int d2[4] = {nu, mu, nu, mu}, s2[4] = {+1, -1, -1, +1};
            const int d3[4] = {mu, nu, mu, nu}, s3[4] = {-1, -1, +1, +1};
            const int d4[4] = {nu, mu, nu, mu}, s4[4] = {-1, +1, +1, -1};
            zc q[9], Q[9] = {0};
            path_product(&g, u, s, c, d1, s1, q); for (int k = 0; k < 9; k++) Q[k] += q[k];
            path_product(&g, u, s, c, d2, s2, q); for (int k = 0; k < 9; k++) Q[k] += q[k];
            path_product(&g, u, s, c, d3, s3, q); for (int k = 0; k < 9; k++) Q[k] += q[k];
            path_product(&g, u, s, c, d4, s4, q); for (int k = 0; k < 9; k++) Q[k] += q[k];
            zc tr = 0;
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) F[p][a + 3 * b] = (Q[a + 3 * b] - conj(Q[b + 3 * a])) / 8.0;
            for (int a = 0; a < 3; a++) tr += F[p][a + 3 * a];
            for (int a = 0; a < 3; a++) F[p][a + 3 * a] -= tr / 3.0;
            if (fmunu) memcpy(fmunu + (s * 6 + p) * 9, F[p], sizeof F[p]);
        }
        zc *A = clov + 72 * s;
        for (int blk = 0; blk < 2; blk++)
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
                zc sum = (i == j) ? 1.0 : 0.0;
                for (p = 0; p < 6; p++)
                    sum += coef * sig[p][2 * blk + i / 3][2 * blk + j / 3] * (I * F[p][(i % 3) + 3 * (j % 3)]);
                A[36 * blk + i + 6 * j] = sum;
            }
    }
}

/* ---- clover-term part of the pseudofermion force (new capability like the clover term itself; groundwork for the device kernel).
 *      S_f = phi^dag (M^dag M)^-1 phi, M = A - kappa H, X = (M^dag M)^-1 phi, Y = M X:  dS = -2 Re[Y^dag dM X].  The hopping part
 *      is orc_wilson_force (it never sees A); the clover part is
 *          -2 Re sum_n Y(n)^dag dA(n) X(n),   dA = kappa csw sum_p sigma_p (x) i dF^_p,   F^_p = (Q_p - Q_p^dag)/8 - trace,
 *      = -2 Re sum_{n,p} (i c / 8) tr[ dQ_p(n) K_p(n) ],   K = Lambda' + Lambda'^dag,  Lambda'= Lambda - tr(Lambda)/3,
 *        Lambda_p(n)[b,a] = sum_{al,be} sigma_p[al,be] X(n)[be,b] conj(Y(n)[al,a]).
 *      Q_p is the sum of the four leaves of orc_clover_build; varying U_rho(m) -> exp(eps A) U_rho(m) inside a leaf
 *      L1 L2 L3 L4 (closed at n) gives tr[A U S K P] for a forward link (P = links before it, S = links after it) and
 *      -tr[A S K P U^dag] for a backward one, so with G the sum of those matrices  dS = -2 Re tr[A (i c / 8) G]:
 *      out_rho(m) += (i c / 8) G_rho(m) in the convention of orc_wilson_force (dS/d eps = -2 Re tr[A out]).  Serial scatter. */
void orc_clover_force(const orc_op *op, zc *const out[4], const zc *const u[4], const zc *X, const zc *Y) {
    geom g = mkgeom(op->dims);
    const int64_t V = g.V;
    const double coef = op->kappa * op->csw;
    zc gam[4][4][4], sig[6][4][4];
    for (int mu = 0; mu < 4; mu++) for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++)
        gam[mu][a][b] = 0.5 * (op->rplusg[mu][a][b] - op->rminusg[mu][a][b]);
    int pl = 0;
    for (int mu = 0; mu < 4; mu++) for (int nu = mu + 1; nu < 4; nu++, pl++)
        for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) {
            zc s = 0;
            for (int k = 0; k < 4; k++) s += gam[mu][a][k] * gam[nu][k][b] - gam[nu][a][k] * gam[mu][k][b];
            sig[pl][a][b] = 0.5 * I * s;
        }
    for (int64_t n = 0; n < V; n++) {
        int c0[4]; site_coords(&g, n, c0);
        int p = 0;
        for (int mu = 0; mu < 4; mu++) for (int nu = mu + 1; nu < 4; nu++, p++) {
            zc K[9];
            {   /* Lambda[b + 3 a] (row b, column a), traceless part, plus its adjoint */
                zc Lm[9], tr = 0;
                for (int b = 0; b < 3; b++) for (int a = 0; a < 3; a++) {
                    zc s = 0;
                    for (int al = 0; al < 4; al++) for (int be = 0; be < 4; be++)
                        s += sig[p][al][be] * X[b + 3 * (n + V * be)] * conj(Y[a + 3 * (n + V * al)]);
                    Lm[b + 3 * a] = s;
                }
                for (int a = 0; a < 3; a++) tr += Lm[a + 3 * a];
                for (int a = 0; a < 3; a++) Lm[a + 3 * a] -= tr / 3.0;
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) K[i + 3 * j] = Lm[i + 3 * j] + conj(Lm[j + 3 * i]);
            }
            const int dirs[4][4] = {{mu, nu, mu, nu}, {nu, mu, nu, mu}, {mu, nu, mu, nu}, {nu, mu, nu, mu}};
            const int sgns[4][4] = {{+1, +1, -1, -1}, {+1, -1, -1, +1}, {-1, -1, +1, +1}, {-1, +1, +1, -1}};
            for (int leaf = 0; leaf < 4; leaf++) {
                zc Lk[4][9], P[5][9], S[5][9];
                int64_t site[4]; int dg[4];
                int c[4] = {c0[0], c0[1], c0[2], c0[3]};
                int64_t s = n;
                for (int k = 0; k < 4; k++) {
                    const int d = dirs[leaf][k];
                    if (sgns[leaf][k] > 0) { site[k] = s; dg[k] = 0; memcpy(Lk[k], u[d] + 9 * s, sizeof Lk[k]); s = hopsite(&g, s, c, d, +1); }
                    else {
                        s = hopsite(&g, s, c, d, -1); site[k] = s; dg[k] = 1;
                        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Lk[k][i + 3 * j] = conj(u[d][9 * s + j + 3 * i]);
                    }
                }
                const zc one[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                memcpy(P[0], one, sizeof one); memcpy(S[4], one, sizeof one);
                for (int k = 0; k < 4; k++) mulx(P[k + 1], P[k], 0, Lk[k], 0);            /* P[k] = L_1 .. L_k */
                for (int k = 3; k >= 0; k--) mulx(S[k], Lk[k], 0, S[k + 1], 0);            /* S[k] = L_{k+1} .. L_4 */
                for (int k = 0; k < 4; k++) {
                    zc t1[9], R[9], G[9];
                    mulx(t1, S[k + 1], 0, K, 0);
                    mulx(R, t1, 0, P[k], 0);                                               /* (links after) K (links before) */
                    if (!dg[k]) mulx(G, Lk[k], 0, R, 0);                                   /* U R */
                    else { mulx(G, R, 0, Lk[k], 0); for (int i = 0; i < 9; i++) G[i] = -G[i]; }     /* -R U^dag */
                    zc *o = out[dirs[leaf][k]] + 9 * site[k];
                    for (int i = 0; i < 9; i++) o[i] += (I * coef / 8.0) * G[i];
                }
            }
        }
    }
}

/* ---- even-odd preconditioned Wilson solve (see lqcd_oracle.h) ---- */
int orc_eo_solve(const orc_op *op, int method, int dagger, zc *x, const zc *const u[4], const zc *b,
                 double eps, int maxsteps, double *resid_sq, double *hist) {
    if (op->csw != 0.0) abort();      /* Mhat below assumes M_ee = M_oo = 1 */
    geom g = mkgeom(op->dims);
    const int64_t V = g.V, n = 12 * V;
    zc *bh = malloc(sizeof(zc) * n), *xe = malloc(sizeof(zc) * n), *t = malloc(sizeof(zc) * n);
    const double kappa = op->kappa;
    /* bhat_e = b_e + kappa H_eo b_o ; xe = even part of the initial guess */
    wilson_apply_p(op, dagger, t, u, b, 0);                  /* reads the odd part of b only */
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < V; s++) {
        int c[4]; site_coords(&g, s, c);
        const int even = ((c[0] + c[1] + c[2] + c[3]) & 1) == 0;
        for (int al = 0; al < 4; al++) for (int a = 0; a < 3; a++) {
            int64_t i = a + 3 * (s + V * al);
            bh[i] = even ? b[i] + kappa * t[i] : 0;
            xe[i] = even ? x[i] : 0;
        }
    }
    orc_op hop = *op;
    int it;
    /* the Krylov routines solve A x = b with A = orc_apply(.., ORC_D ..); for the dagger system hand them Mhat^dag as "D" by
     * swapping the projector tables (dagger swaps (r-g) <-> (r+g), nothing else) */
    if (dagger) { memcpy(hop.rplusg, op->rminusg, sizeof hop.rplusg); memcpy(hop.rminusg, op->rplusg, sizeof hop.rminusg); }
    if (method == 0) it = orc_cgnr(&hop, ORC_WILSON_EO, xe, u, bh, eps, maxsteps, resid_sq, hist);
    else             it = orc_bicgstab(&hop, ORC_WILSON_EO, xe, u, bh, eps, maxsteps, resid_sq, hist);
    /* x_o = b_o + kappa H_oe x_e */
    wilson_apply_p(op, dagger, t, u, xe, 1);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < V; s++) {
        int c[4]; site_coords(&g, s, c);
        const int even = ((c[0] + c[1] + c[2] + c[3]) & 1) == 0;
        for (int al = 0; al < 4; al++) for (int a = 0; a < 3; a++) {
            int64_t i = a + 3 * (s + V * al);
            x[i] = even ? xe[i] : b[i] + kappa * t[i];
        }
    }
    free(bh); free(xe); free(t);
    return it;
}

/* ---- gauge-sector molecular dynamics: the steps of src/md/AbstractMD.jl:78-135 -------------------------
 *  U_update!          U_mu <- exptU(eps*dtau * p_mu) * U_mu                                  (AbstractMD.jl:78-99)
 *  P_update!          p_mu <- p_mu + (-eps*dtau/NC) * Traceless_antihermitian(U_mu * dSdU_mu)  (AbstractMD.jl:101-118),
 *                     plaquette action beta/2 * (P + P^dag) (universe.jl:90-93), S_g = -(beta/NC) sum_plaq Re tr U_p
 *                     (standardHMC.jl:49-50)
 *  P_update_fermion!  p_mu <- p_mu + (-eps*dtau) * Traceless_antihermitian(UdSfdU_mu)        (AbstractMD.jl:120-135)
 *  Momenta are kept as anti-Hermitian traceless 3x3 matrices p = sum_a a_a T_a, T_a = i lambda_a / 2 (upstream stores the
 *  eight real a_a [UPSTREAM-RECALL]); p*p/2 = sum a^2/2 = sum_ij |p_ij|^2.  TA(M) = (M - M^dag)/2 - tr/3.  With these
 *  conventions Hamilton's equations for H = p*p/2 + S_g + S_f give exactly the factors above (beta/(2 NC) for the plaquette
 *  term): the known-answer tests are energy conservation at O(dtau^2) and reversibility, not a reference number.      */
static void ta3(zc *a, const zc *m) {                      /* traceless anti-Hermitian part */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i + 3 * j] = 0.5 * (m[i + 3 * j] - conj(m[j + 3 * i]));
    zc tr = (a[0] + a[4] + a[8]) / 3.0;
    a[0] -= tr; a[4] -= tr; a[8] -= tr;
}
/* exp of a 3x3 matrix: scaling (Frobenius norm <= 1/4), Taylor order 12 by Horner, squaring.  The device code repeats
 * the same operations in the same order (csrc/gauge_md.cu: expm3). */
static void expm3(zc *e, const zc *x0) {
    double n2 = 0;
    for (int i = 0; i < 9; i++) n2 += creal(x0[i]) * creal(x0[i]) + cimag(x0[i]) * cimag(x0[i]);
    double nrm = sqrt(n2), sc = 1.0;
    int sq = 0;
    while (nrm > 0.25) { nrm *= 0.5; sc *= 0.5; sq++; }
    zc x[9], t[9];
    for (int i = 0; i < 9; i++) x[i] = sc * x0[i];
    for (int i = 0; i < 9; i++) e[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int k = 12; k >= 1; k--) {                        /* e = 1 + (x/k) e */
        mm(t, x, e);
        for (int i = 0; i < 9; i++) e[i] = ((i % 4 == 0) ? 1.0 : 0.0) + t[i] / (double)k;
    }
    for (int q = 0; q < sq; q++) { mm(t, e, e); memcpy(e, t, sizeof t); }
}

void orc_md_update_u(const int dims[4], zc *const u[4], const zc *const p[4], double eps) {
    geom g = mkgeom(dims);
    for (int mu = 0; mu < 4; mu++) {
#pragma omp parallel for schedule(static)
        for (int64_t s = 0; s < g.V; s++) {
            zc x[9], e[9], w[9];
            for (int i = 0; i < 9; i++) x[i] = eps * p[mu][9 * s + i];
            expm3(e, x);
            mm(w, e, u[mu] + 9 * s);
            memcpy(u[mu] + 9 * s, w, sizeof w);
        }
    }
}

/* sum of the six staples closing U_mu(n): V such that Re tr[U_mu(n) V] = sum of the six plaquettes through the link */
static void staple_sum(const geom *g, const zc *const u[4], int64_t s, const int c[4], int mu, zc *v) {
    memset(v, 0, 9 * sizeof(zc));
    int w;
    const int64_t smu = nbr(g, s, c, mu, +1, &w);
    int cmu[4] = {c[0], c[1], c[2], c[3]};
    cmu[mu] = (c[mu] + 1) % g->d[mu];
    for (int nu = 0; nu < 4; nu++) {
        if (nu == mu) continue;
        const int64_t snu = nbr(g, s, c, nu, +1, &w), sdn = nbr(g, s, c, nu, -1, &w), smudn = nbr(g, smu, cmu, nu, -1, &w);
        zc a[9], b[9];
        mmd(a, u[nu] + 9 * smu, u[mu] + 9 * snu);         /* U_nu(n+mu) U_mu(n+nu)^dag */
        mmd(b, a, u[nu] + 9 * s);                          /* ... U_nu(n)^dag           */
        for (int i = 0; i < 9; i++) v[i] += b[i];
        /* lower staple: U_nu(n+mu-nu)^dag U_mu(n-nu)^dag U_nu(n-nu) */
        zc t[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {      /* t = U_nu(n+mu-nu)^dag U_mu(n-nu)^dag */
            zc sacc = 0;
            for (int k = 0; k < 3; k++) sacc += conj(u[nu][9 * smudn + k + 3 * i]) * conj(u[mu][9 * sdn + j + 3 * k]);
            t[i + 3 * j] = sacc;
        }
        mm(b, t, u[nu] + 9 * sdn);
        for (int i = 0; i < 9; i++) v[i] += b[i];
    }
}

void orc_md_update_p_gauge(const int dims[4], zc *const p[4], const zc *const u[4], double eps, double beta) {
    geom g = mkgeom(dims);
    const double f = eps * beta / 6.0;                     /* beta / (2 NC) */
    for (int mu = 0; mu < 4; mu++) {
#pragma omp parallel for schedule(static)
        for (int64_t s = 0; s < g.V; s++) {
            int c[4]; site_coords(&g, s, c);
            zc v[9], m[9], a[9];
            staple_sum(&g, u, s, c, mu, v);
            mm(m, u[mu] + 9 * s, v);
            ta3(a, m);
            for (int i = 0; i < 9; i++) p[mu][9 * s + i] -= f * a[i];
        }
    }
}

void orc_md_update_p_force(const int dims[4], zc *const p[4], const zc *const F[4], double eps) {
    geom g = mkgeom(dims);
    for (int mu = 0; mu < 4; mu++) {
#pragma omp parallel for schedule(static)
        for (int64_t s = 0; s < g.V; s++) {
            zc a[9];
            ta3(a, F[mu] + 9 * s);
            for (int i = 0; i < 9; i++) p[mu][9 * s + i] -= eps * a[i];
        }
    }
}

double orc_md_kinetic(const int dims[4], const zc *const p[4]) {
    geom g = mkgeom(dims);
    double sum = 0;
    for (int mu = 0; mu < 4; mu++) {
#pragma omp parallel for reduction(+ : sum) schedule(static)
        for (int64_t i = 0; i < 9 * g.V; i++) sum += creal(p[mu][i]) * creal(p[mu][i]) + cimag(p[mu][i]) * cimag(p[mu][i]);
    }
    return sum;
}

double orc_md_gauge_action(const int dims[4], const zc *const u[4], double beta) {
    geom g = mkgeom(dims);
    return -(beta / 3.0) * orc_plaquette(dims, u) * 18.0 * (double)g.V;
}
