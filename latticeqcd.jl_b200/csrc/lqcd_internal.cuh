// lqcd_internal.cuh -- shared definitions of liblqcd_b200 (sm_100a only).
//
// Device field layout ("AoSoA-32"): sites are numbered lexicographically, x fastest
// (site = x + X*(y + Y*(z + Z*t)) over the LOCAL lattice) and grouped in blocks of 32 consecutive sites
// (one warp).  Inside a block the complex components are stored component-major:
//
//     spinor (ncomp = 12 Wilson, k = 3*alpha + c;  ncomp = 3 staggered, k = c)
//         f[(blk*ncomp + k)*32 + lane]                       double2 = (re, im)
//     links  (4 directions x 9 matrix elements, e = 3*a + b, row a, column b)
//         g[((blk*4 + mu)*9 + e)*32 + lane]                  double2
//
// so that (i) every warp-wide load of one component is one contiguous, 512-byte, 128-bit-per-lane request,
// (ii) all data of a 32-site block is one contiguous record (6 KB spinor, 18 KB links) that a single
// cp.async.bulk can move, and (iii) a block's HBM pages stay together (DRAM page / TLB locality).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/lqcd_b200.h"

#define LQCD_SB 32                 // sites per block (= warp size)
#define LQCD_MAX_RED 8             // doubles reduced per kernel
#define LQCD_MAX_SHIFTS 32

typedef double2 cplx;

struct Geom {
    int X, Y, Z, T;            // local extents
    int V;                     // local volume
    int nblk;                  // V / 32
    // CTA tiling of the block lattice (regular case) -- see make_tiling()
    int regular;               // 1: blocks form a regular 4-d lattice
    int s[4];                  // block shape in sites  (product 32)
    int nb[4];                 // block-lattice extents
    int c[4];                  // CTA tile in blocks     (product = warps per CTA)
    int nt[4];                 // tiles per direction
    int wpc;                   // warps per CTA
    int part[4];               // 1 if direction is partitioned across ranks (neighbour off-rank)
    int gX, gY, gZ, gT;        // global extents
    int o[4];                  // origin of the local lattice in the global one
};

// Solver scalars living in device memory: kernels read alpha/beta from here so that the Krylov loop runs
// without host round trips; `done` turns every later kernel of the loop into a no-op.
struct SolverState {
    double rr;        // current |r|^2
    double pq;        // <p, A p> (or |q|^2 for CGNR)
    double c1;        // CGNR: |A^dag r|^2
    double alpha, beta;
    double eps;
    double are, aim, bre, bim, cre, cim;   // generic complex scalars (BiCGStab)
    double rho_re, rho_im, omega_re, omega_im, alpha_re, alpha_im;
    int done;         // 1 once converged (or broken down)
    int failed;       // 1 if |r|^2 became NaN/Inf (Krylov breakdown)
    int iters;        // iteration at which convergence was detected
    int it;           // running iteration counter
    int maxit;
    int nshift;
    double shift[LQCD_MAX_SHIFTS];
    double zeta[LQCD_MAX_SHIFTS], zeta_old[LQCD_MAX_SHIFTS], alpha_s[LQCD_MAX_SHIFTS], beta_s[LQCD_MAX_SHIFTS];
    double alpha_old, beta_old;
    double red[LQCD_MAX_RED];   // raw reduction results of the last reducing kernel
};

#define LQCD_MAX_RANKS 8
#define LQCD_RED_SLOTS 4
#define LQCD_SPIN_TIMEOUT_DEFAULT_S 20.0          // device-side spin limit: a lost peer becomes LQCD_ERR_COMM, not a hung GPU
                                                    // (override: LQCD_COMM_TIMEOUT_S; ranks may legitimately be seconds apart, e.g.
                                                    //  when one rank's host does CPU work between collective solves)

// In-kernel all-reduce over NVLink peer memory (one process per GPU).  Every rank's reducing kernel
// writes its local totals straight into slot [seq % SLOTS][my rank] of EVERY rank's buffer (peer stores),
// raises a per-(slot, writer) sequence flag with release semantics, then spins on its own local flags until
// all ranks have arrived and sums the contributions in rank order: identical bits on every rank, no NCCL
// launch on the critical path of a Krylov iteration.
struct CommRed {
    int nranks, rank;
    unsigned long long *seq;                      // local device counter of reductions performed
    double *vals[LQCD_MAX_RANKS];                 // vals[r]: rank r's [SLOTS][nranks][LQCD_MAX_RED] array (peer mapped)
    unsigned long long *flags[LQCD_MAX_RANKS];    // flags[r]: rank r's [SLOTS][nranks] sequence flags
    int *err;                                     // local device error word (non-zero = timeout code)
    long long timeout_cycles;
};

// deterministic grid reduction workspace
struct Reduce {
    double *partials;          // [grid * LQCD_MAX_RED]
    unsigned int *ticket;      // last-block-done counter (self-resetting)
    SolverState *st;           // where results / derived scalars go
    double *hist;              // optional per-iteration |r|^2 history (device)
    CommRed cr;                // cr.nranks <= 1: single GPU
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {           // %globaltimer: one clock for all SMs (phase stamps)
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// phase stamps of one application (LQCD_COMM_TIMING=1): [0,1] pack CTAs first start / last end, [2,3] interior tiles, [4,5] face
// tiles, [6] ns all face CTAs spent waiting for the neighbours' flags, [7] the longest such wait
__device__ __forceinline__ void stamp_span(unsigned long long *tm, int cls, unsigned long long t0, unsigned long long t1) {
    atomicMin(tm + 2 * cls, t0);
    atomicMax(tm + 2 * cls + 1, t1);
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

enum FinishOp {
    FIN_STORE = 0,        // st->red[j] = sum_j
    FIN_CG_PQ,            // red0 = Re<p,q>        -> pq, alpha = rr/pq
    FIN_CG_RR,            // red0 = |r_new|^2      -> convergence test, beta, rr, it++
    FIN_CG_INIT,          // red0 = |r0|^2         -> rr, convergence test at step 0
    FIN_CG_PQN,           // red2 = |D p|^2 = <p, D^dag D p> -> pq, alpha   (norm form: no extra read of p)
    FIN_CG_RRN,           // red2 = |r_new|^2 from the fused r -= alpha q epilogue -> convergence test, beta, rr, it++
    FIN_NR_C2,            // red0 = |q|^2          -> alpha = c1/c2
    FIN_NR_RR,            // red0 = |res|^2        -> convergence test, it++
    FIN_NR_C3,            // red0 = |s|^2          -> beta = c3/c1, c1 = c3
    FIN_NR_C1,            // red0 = |q0|^2         -> c1
    FIN_BI_INIT,          // red0 = |r0|^2         -> rr, rho = (rr,0), convergence test at step 0
    FIN_BI_ALPHA,         // red0,1 = <r0,v>       -> alpha = rho / <r0,v>
    FIN_BI_OMEGA,         // red0,1 = <t,s>, red2 = |t|^2 -> omega
    FIN_BI_RR,            // red0 = |r|^2, red1,2 = <r0,r> -> convergence, beta, rho
    FIN_MS_PQ,            // multishift: pq -> alpha, zeta recurrences, alpha_j
    FIN_MS_RR             // multishift: rr_new -> convergence, beta, beta_j
};

struct MSPtrs { cplx *x[LQCD_MAX_SHIFTS]; cplx *p[LQCD_MAX_SHIFTS]; };

struct lqcd_fermion {
    int kind;          // LQCD_WILSON / LQCD_STAGGERED
    int ncomp;         // 12 / 3
    cplx *d;           // device AoSoA-32 data, nblk*ncomp*32 complex
    size_t bytes;
    lqcd_ctx *owner;
};

struct lqcd_ctx {
    int device;
    int rank, nranks;
    int procgrid[4], pcoord[4];
    Geom g;
    cudaStream_t stream, stream2;
    cudaEvent_t ev0, ev1, ev_pack, ev_int, ev_poll[2];
    cplx *gauge;               // AoSoA-32 links
    bool gauge_valid;
    // "two-row" copy of the links for the Dslash kernels (links12.cu): rows 0 and 1 of every SU(3) matrix, the third row is
    // rebuilt in registers as conj(row0 x row1).  Valid for (links12_epoch == gauge_epoch); links12_ok says whether the links
    // passed the unitarity test (|U[2][b] - conj(row0 x row1)[b]| <= 1e-13 everywhere) -- if not, the kernels read the full links.
    cplx *links12; uint64_t links12_epoch; bool links12_ok; double links12_dev; double *links12_scratch;
    // staging
    void *stage; size_t stage_bytes;
    // reductions
    Reduce red;
    SolverState *st_host;      // pinned: [0] init image / blocking reads, [1],[2] async polling slots
    double *hist_dev; int hist_cap;
    // scratch fermions for solvers (allocated lazily per kind)
    std::vector<lqcd_fermion *> scratch[2];
    // L2 flush buffer
    void *flush; size_t flush_bytes;
    cplx *force_buf;           // link-shaped output of the force kernel (allocated on first use, reused every MD step)
    cplx *mom;                 // MD momenta, link-shaped anti-Hermitian traceless matrices (gauge_md.cu)
    bool mom_valid;
    bool force_valid;          // force_buf holds a force (lqcd_fermion_force_xy accumulate / _download)
    uint64_t gauge_epoch;      // bumped by every lqcd_gauge_upload / lqcd_gauge_random
    cplx *clover;              // packed clover term (clover.cu), valid for (clover_epoch, clover_coef)
    uint64_t clover_epoch; double clover_coef;
    cplx *clover_k;            // K_p(n) field of the clover-term force (clover_force.cu), allocated on first use
    uint64_t launches;
    int num_sms;
    mutable std::string err;
    // comm (multi-GPU) -- see comm.cu
    struct CommState *comm;
    // even-odd preconditioned solve -- see wilson_eo.cu
    struct EoState *eo;
    unsigned int *queue;       // tile queue of the persistent-CTA Dslash variants (zero between launches)
    struct MrhsWork *mrhs;     // per-right-hand-side solver states / reduction workspaces of the batched solves -- see mrhs.cu
    struct HostPipe *pipe;     // host-field pipeline (streams, events, two staging buffers) -- see host_pipeline.cu
    int eo_active;             // 1 while lqcd_solve_eo runs the Krylov loop: the solver's operator is Mhat on even half fields
    int stag_even_solve;       // scoped switch (lqcd_set_staggered_even_solve): staggered CG on DdagD runs on even half fields
};

int lqcd_fail(const lqcd_ctx *ctx, int code, const char *fmt, ...);

#define CUDA_TRY(ctx, expr)                                                                             \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return lqcd_fail(ctx, LQCD_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,           \
                             cudaGetErrorString(_e));                                                   \
    } while (0)

#define LQCD_TRY(expr)                                                                                  \
    do {                                                                                                \
        int _s = (expr);                                                                                \
        if (_s != LQCD_OK) return _s;                                                                   \
    } while (0)

static inline int ncomp_of(int kind) { return kind == LQCD_WILSON ? 12 : 3; }

// ---- complex helpers (device) ------------------------------------------------------------------
__device__ __forceinline__ cplx cmake(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {   // conj(a) * b
    return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ void cfma(cplx &acc, cplx a, cplx b) {        // acc += a*b
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfmac(cplx &acc, cplx a, cplx b) {       // acc += conj(a)*b
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ cplx cscale(double s, cplx a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ cplx cmuli(cplx a) { return make_double2(-a.y, a.x); }    // i*a
__device__ __forceinline__ cplx cmulmi(cplx a) { return make_double2(a.y, -a.x); }   // -i*a

__device__ __forceinline__ cplx ldg128(const cplx *p) { return __ldg(p); }
// pull a line towards L2 without holding a register (epilogue operands of the fused Dslash kernels)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- kernels' host-side entry points (one per .cu) ----------------------------------------------
struct DslashFuse {
    // optional fused epilogue: dot_with != nullptr -> reduce  red0/1 = <dot_with, y>, red2 = |y|^2
    const cplx *dot_with;
    int want_norm;             // reduce |y|^2 in red2 even without dot_with
    int finish;                // FinishOp applied to the reduction
    int use_state;             // kernels early-exit when st->done
    double shift;              // y += shift * x  (multi-shift base system: (DdagD + s) )
    const cplx *shift_src;     // field multiplied by `shift` (the input of the first hop of DdagD)
    cplx *axpy_r;              // CG: do not store y; instead r <- r - alpha*y (alpha from SolverState) and reduce |r|^2 in red2
    int cta_off, cta_count;    // single-rank sub-range launch (host_pipeline.cu): CTAs [cta_off, cta_off + cta_count) of the
                               // t-slowest tile order = a slab of t-slices; cta_count = 0 -> whole lattice.  No reductions.
    unsigned int *queue;       // persistent-CTA variants (wilson_kernel.cuh): self-resetting tile queue counter
    int queue_total;           //   number of tiles (pack + Dslash) drawn from it
};

// multi-GPU: halo data consumed INSIDE the Dslash kernel (fused exterior).  CTAs are permuted so that tiles
// without face sites run first and face tiles last; a face CTA spins (ld.acquire.sys + timeout) on the
// neighbours' sequence flags, which by then have normally been raised by the neighbours' pack kernels.
struct HaloIn {
    const cplx *recv[4][2];                     // [mu][0: from lower nbr (my low face), 1: from upper nbr (my high face)]
    const unsigned long long *recv_flag[4][2];
    unsigned long long seq;
    int *err;
    int pfirst[4], plast[4];                    // this rank touches the global low / high boundary in mu
    const int *cta_order;
    int n_interior;
    long long timeout_cycles;
    unsigned long long *timing;                 // LQCD_COMM_TIMING=1: this application's 8 phase stamps (comm.cu), else null
};

struct HaloOut {
    cplx *send[4][2];                    // [mu][0] my LOW face -> lower nbr's "from upper" slot ; [1] my HIGH face -> upper nbr's "from lower" slot
    unsigned long long *send_flag[4][2];
    int cta0[5];                         // first pack CTA of direction mu (prefix over partitioned directions), cta0[4] = npack
    unsigned int *ticket;                // last-pack-CTA detector (self resetting)
    unsigned long long seq;              // application number published in the flags
    int gpu_fence;                       // 1: pack CTAs fence at gpu scope only; the LAST pack CTA's system fence (cumulative) covers them
    unsigned long long *timing;          // LQCD_COMM_TIMING=1: phase stamps of this application, else null
    int spt;                             // face sites per pack thread (>= 1)
};

struct WilsonArgs {
    cplx *out;
    const cplx *in;
    const cplx *gauge;
    const cplx *links12;  // two-row links (G12 kernels), else unused
    Geom g;
    double kappa;
    double bc[4];
    DslashFuse fuse;
    Reduce red;
    HaloIn halo;     // MULTI kernels only
    HaloOut hout;    // MULTI == 2 (self-packing) only
    const cplx *clover;   // CLOVER kernels only: packed clover blocks, 36 complex per site (wilson_kernel.cuh); keep LAST
};

// multi-GPU fermion force: the forward neighbours X(n+mu), Y(n+mu) of the HIGH face sites live on the upper neighbour rank,
// which stores their spin-projected halves (Wilson: P- X | P+ Y, 12 complex per face site; staggered: X | Y, 6) straight
// into this rank's force slot of direction mu (dedicated slots of the comm buffer, one sequence flag per direction).
struct ForceHalo {
    cplx *send[4];                          // LOWER neighbour's slot for direction mu (peer mapped): my low face goes there
    unsigned long long *send_flag[4];
    const cplx *recv[4];                    // my slot: data of the upper neighbour's low face
    const unsigned long long *recv_flag[4];
    unsigned long long seq;                 // force call number published in the flags
    unsigned int *ticket;                   // last-pack-CTA detector (shared with the halo pack: same stream, never concurrent)
    int *err;
    long long timeout_cycles;
    int plast[4];                           // this rank touches the global high boundary in mu (boundary phase on the hop)
    int start[5];                           // pack work items: prefix over partitioned directions of the face sizes
};
int comm_force_halo(lqcd_ctx *ctx, ForceHalo *out);      // comm.cu: bumps the force sequence number

// links12.cu: (re)builds ctx->links12 if stale; *use = 1 if the Dslash kernels may read it (SU(3) links, not switched off)
int ensure_links12(lqcd_ctx *ctx, int *use);
int launch_wilson_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger,
                         const DslashFuse *fuse, cudaStream_t s, const HaloIn *halo = nullptr, const HaloOut *hout = nullptr);
int launch_staggered_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger,
                            const DslashFuse *fuse, cudaStream_t s, const HaloIn *halo = nullptr, const HaloOut *hout = nullptr);
// clover.cu: (re)builds ctx->clover for kappa*csw of `op` if the cached one is stale
int ensure_clover(lqcd_ctx *ctx, const lqcd_op *op);
int launch_wilson_clover(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, int multi, int lh, int grid, int bs, cudaStream_t s);   // wilson_clover.cu
int apply_op(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, const lqcd_fermion *x, int mode,
             lqcd_fermion *tmp, const DslashFuse *fuse_last);

// context.cu: layout conversion of the 32-site blocks [blk0, blk0 + nblk) between the staged host-layout field and the
// device AoSoA-32 field (wing 0), enqueued on stream s
int convert_fermion_range(lqcd_ctx *ctx, int to_device, cplx *dev, cplx *host_stage, int ncomp, int blk0, int nblk, cudaStream_t s);

// blas.cu
int blas_zero(lqcd_ctx *ctx, cplx *x, size_t n);
int blas_copy(lqcd_ctx *ctx, cplx *dst, const cplx *src, size_t n);
// Scratch fermion slots (per kind, allocated on first use).  Who owns what -- callers higher in the list may call the ones below
// while their own slots are live, never the other way round:
//   SCR_FORCE_X / _Y   plain pseudofermion force (X = (DdagD)^-1 eta, Y = D X); Y also serves the rational force after its solve
//   SCR_RATIONAL0 + j  shifted solutions of the rational (RHMC) action, j < LQCD_MAX_SHIFTS
//   SCR_RTERM          spin-diagonal remainder of the Wilson operator with r != 1 (wilson_general_r.cu)
//   SCR_MRHS0 + 16 v + j   work vector v (< 4) of right-hand side j (< 16) of the batched solvers
//   0 .. 2 + LQCD_MAX_SHIFTS   the Krylov loops (solve_impl: 0-5, multi-shift: 0-2 and 3 + j)
enum { SCR_FORCE_X = 8, SCR_FORCE_Y = 9, SCR_RATIONAL0 = 3 + LQCD_MAX_SHIFTS, SCR_RTERM = 3 + 2 * LQCD_MAX_SHIFTS, SCR_MRHS0 = 80 };
int get_scratch(lqcd_ctx *ctx, int kind, int idx, lqcd_fermion **out);
int reduce_grid(const lqcd_ctx *ctx);
