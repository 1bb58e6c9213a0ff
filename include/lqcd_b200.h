/*
 * lqcd_b200.h -- flat C ABI of liblqcd_b200.so: the B200-native (sm_100a) Dirac-operator solve path that
 * LatticeQCD.jl reaches through LatticeDiracOperators.jl / Gaugefields.jl.
 *
 * Nothing like this boundary exists in the reference (it is pure Julia dispatch); each
 * entry point below names the Julia generic function + reference call site (paths relative to
 * /root/reference) whose arithmetic it replaces.  The Julia-side ccall shim is in
 * latticeqcd.jl_b200/julia/LQCDB200.jl and INTEGRATION.md; a ctypes mirror used by the tests and bench.py
 * is in latticeqcd.jl_b200/lqcd_b200/.
 *
 * Conventions
 *   - every function returns an int status (LQCD_OK == 0); never throws, never exits.  On error the
 *     message is retrievable with lqcd_last_error(ctx) (ctx may be NULL for creation failures).
 *     The Julia shim turns non-zero into error(...), matching the reference's convention
 *     (src/system/universe.jl:73-75,130).  Non-convergence after MaxCGstep is an error (LQCD_ERR_NOCONV).
 *   - host arrays are the caller's (Julia GC-owned); they are only read/written during the call.
 *     Device memory belongs to the library behind opaque handles.
 *   - host layouts are the Julia column-major arrays, ComplexF64 = 2 doubles:
 *       links    U[mu]  : [NC, NC, NX+2w, NY+2w, NZ+2w, NT+2w]      (w = ndw, wing width)
 *       Wilson   psi    : [NC, NX+2w, NY+2w, NZ+2w, NT+2w, 4]
 *       stagg.   chi    : [NC, NX+2w, NY+2w, NZ+2w, NT+2w, 1]
 *     with LOCAL extents when the context is one rank of a process grid.
 *   - all exported calls are blocking-on-return unless named *_async.
 *   - one context per process and GPU; multi-GPU = one process per GPU (lqcd_comm_*).
 */
#ifndef LQCD_B200_H
#define LQCD_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LQCD_ABI_VERSION 2        /* 2: + rational-action, multi-RHS, gauge-file, staggered even-site, Z4 entry points (additive) */

enum {
    LQCD_OK = 0,
    LQCD_ERR_ARG = 1,        /* bad argument / unsupported parameter */
    LQCD_ERR_CUDA = 2,       /* CUDA runtime error */
    LQCD_ERR_COMM = 3,       /* inter-GPU communication error / timeout */
    LQCD_ERR_NOCONV = 4,     /* solver hit maxsteps (upstream: error("The CG is not converged!")) */
    LQCD_ERR_NOGPU = 5,      /* no usable CUDA device: there is NO CPU fallback */
    LQCD_ERR_STATE = 6       /* call out of order (e.g. dslash before gauge upload) */
};

enum { LQCD_WILSON = 0, LQCD_STAGGERED = 1 };                 /* params["Dirac_operator"], universe.jl:106-113 */
enum { LQCD_OP_D = 0, LQCD_OP_DDAG = 1, LQCD_OP_DDAGD = 2 };  /* D, D', DdagD (measure_Pion_correlator.jl:379) */
enum {
    LQCD_SOLVER_CG = 0,        /* Hermitian A (DdagD): upstream cg           */
    LQCD_SOLVER_CGNR = 1,      /* A = D: upstream "bicg" (really CGNR)      (App. C.4)            */
    LQCD_SOLVER_BICGSTAB = 2   /* A = D: params["method_CG"] = "bicgstab"                          */
};

typedef struct lqcd_ctx lqcd_ctx;
typedef struct lqcd_fermion lqcd_fermion;

/* Operator descriptor = the params Dict of Dirac_operator(U, x, params), universe.jl:103-137 */
typedef struct {
    int kind;          /* LQCD_WILSON | LQCD_STAGGERED                          params["Dirac_operator"] */
    double kappa;      /* Wilson hopping parameter                              params["κ"]  (universe.jl:114) */
    double r;          /* Wilson parameter (r != 1: slower two-kernel route,    params["r"]  (universe.jl:115)
                          single rank, no force / multi-shift / even-odd) */
    double mass;       /* staggered mass                                        params["mass"] (universe.jl:109) */
    double csw;        /* clover coefficient (0 = none; Clover is not reachable from run_LQCD) */
    double bc[4];      /* fermion boundary phases, +-1                          params["boundarycondition"] (universe.jl:135) */
} lqcd_op;

/* ---- context ------------------------------------------------------------------------------------ */
/* global_dims = (NX,NY,NZ,NT); procgrid = ranks per direction (product = nranks; mirrors the reference's
 * PEs[4], src/mpi/mpimodule.jl:9-13); device = CUDA ordinal for this rank.  Replaces the geometry part of
 * Initialize_Gaugefields (universe.jl:41-49). */
int lqcd_ctx_create(const int global_dims[4], const int procgrid[4], int rank, int device, lqcd_ctx **out);
int lqcd_ctx_destroy(lqcd_ctx *ctx);
const char *lqcd_last_error(const lqcd_ctx *ctx);
int lqcd_abi_version(void);
int lqcd_local_dims(const lqcd_ctx *ctx, int local_dims[4], int origin[4]);
int lqcd_synchronize(lqcd_ctx *ctx);
/* pin / unpin a caller-owned host buffer so uploads run at PCIe speed (optional) */
int lqcd_host_register(lqcd_ctx *ctx, void *ptr, size_t bytes);
int lqcd_host_unregister(lqcd_ctx *ctx, void *ptr);

/* ---- link field (Gaugefields.jl container -> device mirror) ------------------------------------ */
/* Upload the 4 link arrays; the natural call point is Dirac_operator(U,x,params) (universe.jl:137) and the
 * rebinding D(U) (measure_Pion_correlator.jl:338) because U mutates every MD step (AbstractMD.jl:89-93). */
int lqcd_gauge_upload(lqcd_ctx *ctx, const double *const U_mu[4], int nc, int ndw);
int lqcd_gauge_download(lqcd_ctx *ctx, double *const U_mu[4], int nc, int ndw);
/* synthetic inputs generated on the device: warm_eps < 0 -> hot (Haar) links,
 * warm_eps >= 0 -> exp(i eps H).  counter-based: identical fields for any process grid. */
int lqcd_gauge_random(lqcd_ctx *ctx, uint64_t seed, double warm_eps);
/* average plaquette of the device links (QCDMeasurements Plaquette; used to pin the loader/layout) */
int lqcd_gauge_plaquette(lqcd_ctx *ctx, double *plaq);

/* ---- pseudofermion fields: Initialize_pseudofermion_fields (universe.jl:107,112) ---------------- */
int lqcd_fermion_alloc(lqcd_ctx *ctx, int kind, lqcd_fermion **out);
int lqcd_fermion_free(lqcd_ctx *ctx, lqcd_fermion *f);
int lqcd_fermion_upload(lqcd_ctx *ctx, lqcd_fermion *f, const double *host, int ndw);
int lqcd_fermion_download(lqcd_ctx *ctx, const lqcd_fermion *f, double *host, int ndw);
int lqcd_fermion_zero(lqcd_ctx *ctx, lqcd_fermion *f);                               /* clear_fermion! */
int lqcd_fermion_copy(lqcd_ctx *ctx, lqcd_fermion *dst, const lqcd_fermion *src);    /* substitute_fermion! */
int lqcd_fermion_gaussian(lqcd_ctx *ctx, lqcd_fermion *f, uint64_t seed);            /* gauss_distribution_fermion!, sigma^2=1/2 */
int lqcd_fermion_z4(lqcd_ctx *ctx, lqcd_fermion *f, uint64_t seed);                  /* Z4_distribution_fermi! (measure_chiral_condensate.jl:181) */
int lqcd_fermion_point_source(lqcd_ctx *ctx, lqcd_fermion *f, const int site[4], int color, int spin);
/* zero the sites of the other parity: keep (x+y+z+t) % 2 == parity (global coordinates).  Staggered Nf = 4 keeps its
 * pseudofermions on even sites only (D^dag D does not mix parities, so solves stay on that parity). */
int lqcd_fermion_mask_parity(lqcd_ctx *ctx, lqcd_fermion *f, int parity);
                                                                                     /* setindex_global! (measure_Pion_correlator.jl:376) */

/* ---- BLAS-1 on fermion fields ------------------------------------------- */
int lqcd_blas_axpy(lqcd_ctx *ctx, double a_re, double a_im, const lqcd_fermion *x, lqcd_fermion *y);   /* add!(y, a, x)      */
int lqcd_blas_xpby(lqcd_ctx *ctx, const lqcd_fermion *x, double b_re, double b_im, lqcd_fermion *y);   /* add!(b, y, 1, x)   */
int lqcd_blas_scale(lqcd_ctx *ctx, double a_re, double a_im, lqcd_fermion *x);
int lqcd_blas_dot(lqcd_ctx *ctx, const lqcd_fermion *a, const lqcd_fermion *b, double out[2]);         /* dot(a,b)=sum conj(a) b (standardHMC.jl:54) */
int lqcd_blas_norm2(lqcd_ctx *ctx, const lqcd_fermion *a, double *out);

/* ---- operator application: LinearAlgebra.mul!(y, D, x), mul!(y, D', x), mul!(y, DdagD, x) -------- */
int lqcd_dslash(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *x, int mode);

/* ---- Wilson-clover term (op->csw != 0; NEW capability: BASELINE.json configs[3]; the surveyed wrapper parses
 *      Clover_coefficient at src/system/parameter_structs.jl:125 but cannot reach a clover operator, universe.jl:106-131).
 *      A(n) = 1 + kappa*csw*sum_{mu<nu} sigma_mu_nu (x) i F^_mu_nu(n), four-leaf F^ as in
 *      src/measurements/unusedfiles/measure_topological_charge.jl:299-309.  The term is (re)built lazily from the device
 *      links by the first lqcd_dslash / lqcd_solve that uses csw != 0 after a gauge upload; this call forces the build and,
 *      when out != NULL, returns the dense blocks out[((site*2 + b)*36 + i + 6*j)] (complex re,im; b = chirality block,
 *      i = 3*spin_in_block + colour).  Multi-rank: all ranks must have finished lqcd_gauge_upload (host barrier) first. */
int lqcd_clover_term(lqcd_ctx *ctx, const lqcd_op *op, double *out);

/* ---- solvers: solve_DinvX!(y, A, x) (measure_Pion_correlator.jl:399, measure_chiral_condensate.jl:182;
 *      inside calc_UdSfdU! AbstractMD.jl:129 and evaluate_FermiAction standardHMC.jl:69-71) ---------
 * y is initial guess and result.  eps is compared with the ABSOLUTE SQUARED residual (params["eps_CG"],
 * universe.jl:132, default 1e-19 parameter_structs.jl:174); maxsteps = params["MaxCGstep"] (universe.jl:134).
 * method/target: CG needs target LQCD_OP_DDAGD; CGNR and BICGSTAB need LQCD_OP_D or LQCD_OP_DDAG.
 * hist (nullable, maxsteps+1 doubles) receives |r|^2 per step (verbose_level 3 prints it upstream). */
int lqcd_solve(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *x, int method, int target,
               double eps, int maxsteps, int *iters, double *resid_sq, double *hist);
/* multi-shift CG: (DdagD + shifts[j]) ys[j] = x, zero initial guess, shifts[0] smallest (RHMC, App. C.5) */
int lqcd_multishift_cg(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *const ys[], const lqcd_fermion *x,
                       const double *shifts, int nshift, double eps, int maxsteps, int *iters, double *resid_sq);

/* even-odd (Schur complement) preconditioned solve of D y = x (target LQCD_OP_D) or D^dag y = x (LQCD_OP_DDAG) for the
 * Wilson operator without clover, r = 1, single rank, even extents (BASELINE.json configs[1]; NEW capability: the wrapper's
 * `isevenodd` is a heatbath flag only, src/updates/AbstractUpdate.jl:97):
 *     xhat_e = x_e + kappa H_eo x_o;   (1 - kappa^2 H_eo H_oe) y_e = xhat_e;   y_o = x_o + kappa H_oe y_e
 * on checkerboarded half-lattice fields.  method = LQCD_SOLVER_CGNR (upstream "bicg") or LQCD_SOLVER_BICGSTAB, applied to
 * the preconditioned operator; the even part of y is the initial guess; eps / maxsteps / hist as in lqcd_solve -- the
 * preconditioned residual equals the true residual |x - D y|^2 of the full system. */
int lqcd_solve_eo(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *x, int method, int target,
                  double eps, int maxsteps, int *iters, double *resid_sq, double *hist);

/* ---- mul!(y, D, x) on HOST fields (the call the Julia shim makes for the reference's CPU pseudofermion types,
 *      SURVEY.md 8b "Selection" option 1): upload + Dslash + download as one operation, pipelined over slabs of t-slices so
 *      that the H2D copy, the kernels and the D2H copy overlap (both PCIe directions busy at once).  x_host / y_host: host
 *      layout of lqcd_fermion_upload / _download with wing ndw (pin them with lqcd_host_register for asynchronous copies);
 *      x, y: device fields of the operator's kind, left holding the source and the result.  Falls back to the three-call
 *      sequence where the pipeline does not apply (several ranks, wings, irregular tiling). */
int lqcd_dslash_host(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, lqcd_fermion *x, double *y_host, const double *x_host,
                     int mode, int ndw);

/* ---- staggered even-site systems on half fields (csrc/staggered_eo.cu) ------------------------------------------------------------
 *      Staggered Nf = 4 (the reference's default staggered setup, test/test_staggered.toml) keeps its pseudofermion on even sites;
 *      DdagD = m^2 - Dh^2 does not couple the parities, so its solves are A_ee x_e = b_e with A_ee = m^2 - Dh_eo Dh_oe.
 *      lqcd_solve_staggered_even      CG on A_ee over checkerboarded half fields: odd sites of b ignored, even part of y = initial
 *                                     guess, y returns with zero odd sites.  Single rank, even extents.
 *      lqcd_set_staggered_even_solve  scoped switch (default 0): while on, lqcd_solve(CG, DdagD) on a staggered operator -- also
 *                                     inside lqcd_fermion_force and lqcd_md_trajectory -- takes that path; the caller guarantees
 *                                     even-site sources. */
int lqcd_solve_staggered_even(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *b, double eps, int maxsteps,
                              int *iters, double *resid_sq);
int lqcd_set_staggered_even_solve(lqcd_ctx *ctx, int on);

/* ---- gauge configurations in the reference's file formats (csrc/gauge_io.cu) ------------------------------
 *      `initial = "<file>"` + loadU_format (src/system/universe.jl:58-77: ILDG :62-65, load_BridgeText! :66-68) and saveU_format
 *      (src/system/lqcd.jl:226-247: save_binarydata "ILDG", save_textdata "BridgeText").  format: LQCD_IO_ILDG = one LIME record
 *      "ildg-binary-data" of big-endian float64, LQCD_IO_BRIDGETEXT = one float64 per line as Julia prints it; both site-major
 *      (x fastest), per site mu, row, column, (re, im).  Files written here are byte-identical to the reference's own fixtures
 *      (test/confs_HMC_L04040404_beta5.7_Wilson_kappa0.141139/conf_00000100.ildg and .ildg.txt) when given the same links.
 *      lqcd_io_read_gauge / _write_gauge   HOST only (no GPU, no context; errors through lqcd_last_error(NULL)): file <-> the four
 *                                          Julia-layout arrays of lqcd_gauge_upload, wing ndw, any NC
 *      lqcd_gauge_load / lqcd_gauge_save   file <-> device links (NC = 3): every rank moves only the rows of its own block; byte
 *                                          swap and the site-major <-> AoSoA-32 transposition run in one kernel.  ILDG save across
 *                                          ranks: the rank owning the lattice origin creates the file -- call it there first,
 *                                          then (host barrier) on the others.  BridgeText save: single rank. */
#define LQCD_IO_ILDG 0
#define LQCD_IO_BRIDGETEXT 1
int lqcd_io_read_gauge(const char *path, int format, const int dims[4], int nc, double *const U_mu[4], int ndw);
int lqcd_io_write_gauge(const char *path, int format, const int dims[4], int nc, const double *const U_mu[4], int ndw);
int lqcd_gauge_load(lqcd_ctx *ctx, const char *path, int format);
int lqcd_gauge_save(lqcd_ctx *ctx, const char *path, int format);

/* ---- several right-hand sides in lock step ----------------------------------------------------------
 *      The reference's measurement solves are loops of independent solves against the same links: the NC*Nspinor point sources
 *      of calc_quark_propagators_point_source (src/measurements/unusedfiles/measure_Pion_correlator.jl:333-349, solve_DinvX! at
 *      :399) and the Nr noise vectors of the chiral condensate (measure_chiral_condensate.jl:176-182).  For the staggered
 *      operator every link fetched from HBM serves a group of right-hand sides that advance in lock step (csrc/mrhs.cu); each
 *      right-hand side keeps its own Krylov scalars, stopping test (reference rule, eps absolute and squared) and iteration count.
 *      lqcd_dslash_multi   ys[j] = op(mode) xs[j], j < nrhs <= 16: bit-identical to nrhs calls of lqcd_dslash
 *      lqcd_solve_multi    op(target) ys[j] = bs[j], ys[j] is the initial guess; method CGNR (upstream "bicg", target D / D^dag:
 *                          per right-hand side bit-identical to lqcd_solve) or CG (target DdagD); BICGSTAB runs the right-hand
 *                          sides one after the other.  iters / resid_sq: arrays of nrhs (nullable).  LQCD_ERR_NOCONV if any
 *                          right-hand side did not converge (the arrays are filled for all of them).
 *      Wilson (any variant: a Wilson multi-RHS kernel measured slower than the single-RHS kernel and was removed), several ranks:
 *      the same entry points take the single-RHS path, one right-hand side after the other, with the same results. */
int lqcd_dslash_multi(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *const ys[], const lqcd_fermion *const xs[], int nrhs, int mode);
int lqcd_solve_multi(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *const ys[], const lqcd_fermion *const bs[], int nrhs,
                     int method, int target, double eps, int maxsteps, int *iters, double *resid_sq);

/* ---- fermion force: calc_UdSfdU!(UdSfdU, fermi_action, U, eta) (AbstractMD.jl:129) ---------------
 * Given eta: X = (DdagD)^-1 eta by CG, Y = D X, then the 4 link-shaped outer-product fields are written
 * to the host arrays out_mu (same layout as the links, wing width ndw; the halo is zero-filled).
 * x_inout (nullable) is X (initial guess / result).  Wilson-clover (op->csw != 0): hopping part + clover-term derivative
 * (csrc/clover_force.cu; single rank). */
int lqcd_fermion_force(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, lqcd_fermion *x_inout,
                       double eps, int maxsteps, double *const out_mu[4], int ndw, int *iters, double *action);
/* The bilinear part alone, for rational actions (RHMC, staggered Nf not in {4, 8}: README.md:132, test/test_Nf2.toml):
 * calc_UdSfdU! = sum_j alpha_j force(X_j, Y_j) with X_j = (DdagD + beta_j)^-1 phi from lqcd_multishift_cg and Y_j = D X_j.
 * F <- coef * force(X, Y) (accumulate = 0) or F += coef * force(X, Y) in a link-shaped DEVICE buffer owned by the context;
 * lqcd_fermion_force_download copies it to the four host arrays (link layout, wing ndw).  Collective across ranks. */
int lqcd_fermion_force_xy(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *X, const lqcd_fermion *Y, double coef, int accumulate);
int lqcd_fermion_force_download(lqcd_ctx *ctx, double *const out_mu[4], int ndw);
/* calc_UdSfdU! for the RHMC action in one call: multi-shift CG, Y_j = D X_j, sum_j alpha[j] force(X_j, Y_j) on the device, then the
 * copy to out_mu as in lqcd_fermion_force. */
int lqcd_fermion_force_rational(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, const double *alpha, const double *shifts,
                                int nshift, double eps, int maxsteps, double *const out_mu[4], int ndw, int *iters);
/* y = alpha0 x + sum_j alpha[j] (DdagD + shifts[j])^-1 x with ONE multi-shift CG (shifts[0] the smallest): the RHMC heat bath
 * eta = (DdagD)^{Nf/16} xi of sample_pseudofermions! (standardMD.jl:96) and, through dot_out = Re<x, y> (nullable), the action
 * eta^dag (DdagD)^{-Nf/8} eta of evaluate_FermiAction (standardHMC.jl:69-71), as single calls.  Collective across ranks. */
int lqcd_rational_apply(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *x, double alpha0, const double *alpha,
                        const double *shifts, int nshift, double eps, int maxsteps, int *iters, double *dot_out);

/* ---- gauge-sector molecular dynamics on the device: the steps of src/md/AbstractMD.jl:78-135 and the
 *      leapfrog integrators of src/md/standardMD.jl:125-165 for NC = 3 and the plaquette action beta/2*(P + P^dag)
 *      (src/system/universe.jl:85-93), so that links, momenta and pseudofermions stay in HBM for a whole trajectory.
 *      Momenta: link-shaped device field of anti-Hermitian traceless matrices p = sum_a a_a i lambda_a/2 (host layout = link
 *      layout).  Single and multi rank (staples / plaquettes read the neighbour ranks' peer-mapped links; collective calls).
 *      lqcd_md_momenta_gaussian   gauss_distribution!(p)   (standardMD.jl:86): a_a ~ N(0,1), counter-based generator
 *      lqcd_md_kinetic            md.p * md.p / 2          (standardHMC.jl:47)
 *      lqcd_md_gauge_action       -evaluate_GaugeAction(gauge_action, U)/NC = -(beta/NC) sum_plaq Re tr U_p (standardHMC.jl:49-50)
 *      lqcd_md_update_u           U_update!:  U_mu <- exp(eps p_mu) U_mu                      (eps = the reference's eps*dtau)
 *      lqcd_md_update_p           P_update!:  p_mu -= eps beta/(2 NC) TA(U_mu * staples)
 *      lqcd_md_update_p_fermion   P_update_fermion!: p_mu -= eps TA(UdSfdU_mu), the CG + force run inside (zero initial guess)
 *      lqcd_md_trajectory         runMD!: mdsteps leapfrog steps of size dtau; nsw = 0 runMD_QPQ!, nsw > 0 (even) runMD_QPQ_sw!
 *                                 (Sexton-Weingarten: nsw gauge sub-steps around one fermion force); op = NULL: quenched. */
int lqcd_md_momenta_gaussian(lqcd_ctx *ctx, uint64_t seed);
int lqcd_md_momenta_upload(lqcd_ctx *ctx, const double *const P_mu[4], int ndw);
int lqcd_md_momenta_download(lqcd_ctx *ctx, double *const P_mu[4], int ndw);
int lqcd_md_kinetic(lqcd_ctx *ctx, double *out);
int lqcd_md_gauge_action(lqcd_ctx *ctx, double beta, double *out);
int lqcd_md_update_u(lqcd_ctx *ctx, double eps);
int lqcd_md_update_p(lqcd_ctx *ctx, double eps, double beta);
int lqcd_md_update_p_fermion(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, double eps, double cg_eps, int cg_maxsteps, int *iters);
int lqcd_md_trajectory(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, double beta, double dtau, int mdsteps, int nsw,
                       double cg_eps, int cg_maxsteps, long long *cg_iters_total);
/* The same two with the RHMC pseudofermion action eta^dag [alpha0 + sum_j alpha[j] / (DdagD + shifts[j])] eta (staggered Nf not in
 * {4, 8}: README.md:132, test/test_Nf2.toml, test_Nf3.toml): every fermion force is one multi-shift CG and nshift accumulated outer
 * products, nothing leaves the device during the trajectory. */
int lqcd_md_update_p_fermion_rational(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, const double *alpha, const double *shifts,
                                      int nshift, double eps, double cg_eps, int cg_maxsteps, int *iters);
int lqcd_md_trajectory_rational(lqcd_ctx *ctx, const lqcd_op *op, const lqcd_fermion *eta, const double *alpha, const double *shifts,
                                int nshift, double beta, double dtau, int mdsteps, int nsw, double cg_eps, int cg_maxsteps,
                                long long *cg_iters_total);

/* ---- multi-GPU plumbing (one process per GPU; handles are exchanged by the host: MPI.jl Allgather in
 *      Julia, torch.distributed.all_gather in the Python mirror) ---------------------------------- */
#define LQCD_IPC_HANDLE_BYTES 256
int lqcd_comm_export(lqcd_ctx *ctx, void *handle_out /* LQCD_IPC_HANDLE_BYTES */);
int lqcd_comm_connect(lqcd_ctx *ctx, const void *all_handles /* nranks * LQCD_IPC_HANDLE_BYTES, rank order */);
/* pure geometry (no GPU): local extents, origin and +-mu neighbour ranks of `rank` in the process grid;
 * the PEs bookkeeping of src/mpi/mpimodule.jl:9-38 restated for the C side. */
int lqcd_decompose(const int global_dims[4], const int procgrid[4], int rank, int local_dims[4], int origin[4],
                   int nbr_lo[4], int nbr_hi[4]);

/* ---- instrumentation --------------------------------------------------------------------------- */
/* number of kernels this library launched since context creation (bench.py's gpu_launches) */
int lqcd_launch_count(const lqcd_ctx *ctx, uint64_t *count);
/* time `reps` back-to-back applications of one operator with CUDA events on the library's stream;
 * flush_l2 != 0 writes a >L2 scratch buffer between applications (outside the event brackets).
 * Returns the mean per-application milliseconds of the dslash kernels only. */
int lqcd_time_dslash(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *x, int mode,
                     int reps, int flush_l2, double *ms_mean, double *ms_min);
/* expose the CUDA stream (cudaStream_t as void*) so the host can bracket work with its own events */
int lqcd_stream(lqcd_ctx *ctx, void **stream_out);

#ifdef __cplusplus
}
#endif
#endif /* LQCD_B200_H */
