// comm.cu -- multi-GPU domain decomposition: one process per GPU, halos and reductions over NVLink peer memory.
//
// The reference only carries a 4-entry process grid (PEs, src/mpi/mpimodule.jl:9-13, src/mpirun.jl:17-19) that
// is never wired to the fields; upstream's *_mpi field types exchange "wing" halos with
// MPI point-to-point.  B200-native replacement: every rank owns one device buffer (halo slots + flags +
// all-reduce slots) that all other ranks map through CUDA IPC; the data path is entirely in-kernel:
//
//   pack kernel      spin-projects the face spinors (Wilson: 12 -> 6 complex per site; for the backward hop the
//                    sender also applies U^dag so no remote links are ever needed) and STORES them directly into
//                    the neighbour's halo slot over NVLink, then raises a sequence flag (st.release.sys);
//   interior kernel  the normal Dslash kernel with off-rank hops masked (Geom.part) -- runs while the
//                    stores are in flight;
//   exterior kernel  spins on the local flag (ld.acquire.sys), applies U (forward hop), reconstructs and
//                    adds the halo contributions to the face sites;
//   reductions       in-kernel all-reduce (reduce.cuh / CommRed).
//
// Protocol and why it is safe (k = application number = HaloOut.seq, slot = k & 1):
//   producer (pack CTAs of rank A, application k)  : peer stores of the face data into B's slot[k&1]; bar.sync; thread 0:
//       fence (sys) + ticket; the LAST pack CTA: fence.sys, then st.release.sys flag_B[mu][side][k&1] = k.
//   consumer (face CTAs of rank B, application k)  : thread 0 spins on ld.acquire.sys flag >= k (clock64 timeout ->
//       error word, never a hang); bar.sync; halo data is read with ld.global.cg (never cached in L1).
//   write-after-read on a slot: A's pack(k+2) reuses slot[k&1].  It runs after A's Dslash(k+1) finished (stream order),
//       whose face CTAs waited for B's pack(k+1) flag, which B raises only inside its application k+1, i.e. after B's
//       Dslash(k) -- the reader of slot[k&1] -- has completed.  Hence two slots suffice.
//   skipped applications: kernels launched after a solver converged return at once on every rank alike (the converged
//       flag derives from all-reduced, bit-identical values), so no rank ever waits for a flag that is not raised; flags are
//       monotonic sequence numbers, a later application simply publishes a larger one.
//   deadlock freedom: self-packing kernels dispatch their pack CTAs first (lowest blockIdx) and their face CTAs last, pack
//       CTAs never wait; with the separate pack kernel the pack stream has the highest priority so that its CTAs are
//       dispatched before the remaining Dslash CTAs.  A rank's application k only needs its neighbours' pack(k), which
//       needs nothing from this rank's application k.
//   all-reduce (reduce.cuh): slot = seq % 4 with a per-(slot, writer) sequence flag; a rank can be at most one reduction
//       ahead of any other (every reduction waits for all ranks), so a slot is never overwritten before it was read.
// All spin loops carry a clock64 timeout (LQCD_COMM_TIMEOUT_S, default 20 s) that turns a lost peer into LQCD_ERR_COMM
// instead of a hung GPU.  torch.distributed / MPI is used by the HOST only to all-gather the 256-byte handles
// (lqcd_comm_export -> lqcd_comm_connect) and to barrier.
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "wilson_spin.cuh"
#include "site_map.cuh"
#include "halo_pack.cuh"
#include <unistd.h>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct HandleBlob {
    uint32_t magic;
    int32_t rank;
    int64_t pid;
    int32_t device;
    int32_t pad;
    uint64_t raw_ptr;
    uint64_t bytes;
    cudaIpcMemHandle_t ipc;
    // link array of the rank (read by peers' setup kernels that need links beyond the local lattice: clover leaves)
    uint32_t has_gauge; uint32_t pad2;
    uint64_t gauge_raw;
    cudaIpcMemHandle_t gauge_ipc;
};
static_assert(sizeof(HandleBlob) <= LQCD_IPC_HANDLE_BYTES, "handle blob too large");
#define LQCD_HANDLE_MAGIC 0x4c514344u

struct CommState {
    char *base;                       // my buffer
    size_t bytes;
    char *peer[LQCD_MAX_RANKS];       // every rank's buffer mapped here (peer[rank] == base)
    bool opened[LQCD_MAX_RANKS];
    bool connected;
    size_t off_red_flags, off_red_vals, off_halo_flags, off_seq, off_err, off_ticket, off_halo;
    size_t off_force_flags, force_off[4];   // fermion-force halo (force.cu): one flag + one 12-complex-per-face-site slot per direction
    unsigned long long force_seq;
    size_t halo_slot_bytes[4];        // per direction, one (side, slot) buffer
    size_t halo_off[4][2][2];         // [mu][side: 0 = arrives from lower nbr, 1 = from upper nbr][slot]
    int face[4];                      // face sites per direction
    int nbr[4][2];                    // neighbour ranks [mu][0 lower, 1 upper]
    unsigned long long halo_seq;
    int *bsites;                      // device table of boundary sites (ascending site index)
    int nbsites;
    int *cta_order;                   // Dslash CTA permutation: tiles without face sites first, face tiles last
    int n_interior;
    const cplx *gauge_peer[LQCD_MAX_RANKS];   // every rank's link array (nullptr where the mapping failed)
    bool gauge_opened[LQCD_MAX_RANKS];
};

static int rank_of(const lqcd_ctx *ctx, const int pc[4]) {
    return pc[0] + ctx->procgrid[0] * (pc[1] + ctx->procgrid[1] * (pc[2] + ctx->procgrid[2] * pc[3]));
}

static int comm_alloc(lqcd_ctx *ctx) {
    if (ctx->comm) return LQCD_OK;
    if (ctx->nranks > LQCD_MAX_RANKS) return lqcd_fail(ctx, LQCD_ERR_ARG, "at most %d ranks", LQCD_MAX_RANKS);
    CommState *c = new CommState();
    memset(c, 0, sizeof *c);
    const Geom &g = ctx->g;
    const int d[4] = {g.X, g.Y, g.Z, g.T};
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 255) / 256 * 256; return o; };
    c->off_red_flags = take(sizeof(unsigned long long) * LQCD_RED_SLOTS * LQCD_MAX_RANKS);
    c->off_red_vals = take(sizeof(double) * LQCD_RED_SLOTS * LQCD_MAX_RANKS * LQCD_MAX_RED);
    c->off_halo_flags = take(sizeof(unsigned long long) * 4 * 2 * 2);
    c->off_seq = take(sizeof(unsigned long long));
    c->off_err = take(sizeof(int));
    c->off_ticket = take(sizeof(unsigned int));
    c->off_force_flags = take(sizeof(unsigned long long) * 4);
    for (int mu = 0; mu < 4; mu++) {
        c->face[mu] = g.V / d[mu];
        int pc[4] = {ctx->pcoord[0], ctx->pcoord[1], ctx->pcoord[2], ctx->pcoord[3]};
        pc[mu] = (ctx->pcoord[mu] + ctx->procgrid[mu] - 1) % ctx->procgrid[mu]; c->nbr[mu][0] = rank_of(ctx, pc);
        pc[mu] = (ctx->pcoord[mu] + 1) % ctx->procgrid[mu];                      c->nbr[mu][1] = rank_of(ctx, pc);
        if (!g.part[mu]) continue;
        size_t fb = ((size_t)(c->face[mu] + 31) / 32) * 32 * 6 * sizeof(cplx);
        c->halo_slot_bytes[mu] = fb;
        for (int side = 0; side < 2; side++)
            for (int slot = 0; slot < 2; slot++) c->halo_off[mu][side][slot] = take(fb);
        c->force_off[mu] = take(2 * fb);
    }
    c->bytes = off;
    cudaError_t e = cudaMalloc(&c->base, c->bytes);
    if (e != cudaSuccess) { delete c; return lqcd_fail(ctx, LQCD_ERR_CUDA, "cudaMalloc(comm %zu) -> %s", off, cudaGetErrorString(e)); }
    e = cudaMemset(c->base, 0, c->bytes);
    if (e != cudaSuccess) { cudaFree(c->base); delete c; return lqcd_fail(ctx, LQCD_ERR_CUDA, "cudaMemset(comm) -> %s", cudaGetErrorString(e)); }
    {   // boundary-site table: sites with a coordinate on a face of a partitioned direction
        std::vector<int> bs;
        for (int s = 0; s < g.V; s++) {
            int r = s, cc[4];
            for (int i = 0; i < 4; i++) { cc[i] = r % d[i]; r /= d[i]; }
            bool b = false;
            for (int i = 0; i < 4; i++) if (g.part[i] && (cc[i] == 0 || cc[i] == d[i] - 1)) b = true;
            if (b) bs.push_back(s);
        }
        c->nbsites = (int)bs.size();
        e = cudaMalloc(&c->bsites, sizeof(int) * (bs.size() + 1));
        if (e == cudaSuccess) e = cudaMemcpy(c->bsites, bs.data(), sizeof(int) * bs.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(c->base); delete c; return lqcd_fail(ctx, LQCD_ERR_CUDA, "boundary table -> %s", cudaGetErrorString(e)); }
    }
    {   // CTA order for the fused-halo Dslash kernels
        const int ncta = (g.nblk + g.wpc - 1) / g.wpc;
        std::vector<int> inner, face;
        for (int cta = 0; cta < ncta; cta++) {
            bool b = false;
            for (int w = 0; w < g.wpc && !b; w++) {
                const int blk = block_of_warp(g, cta, w);
                if (blk >= g.nblk) continue;
                for (int l = 0; l < 32 && !b; l++) {
                    int r = blk * 32 + l, cc[4];
                    for (int i = 0; i < 4; i++) { cc[i] = r % d[i]; r /= d[i]; }
                    for (int i = 0; i < 4; i++) if (g.part[i] && (cc[i] == 0 || cc[i] == d[i] - 1)) b = true;
                }
            }
            (b ? face : inner).push_back(cta);
        }
        c->n_interior = (int)inner.size();
        inner.insert(inner.end(), face.begin(), face.end());
        e = cudaMalloc(&c->cta_order, sizeof(int) * (inner.size() + 1));
        if (e == cudaSuccess) e = cudaMemcpy(c->cta_order, inner.data(), sizeof(int) * inner.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(c->base); cudaFree(c->bsites); delete c; return lqcd_fail(ctx, LQCD_ERR_CUDA, "cta order table -> %s", cudaGetErrorString(e)); }
    }
    c->peer[ctx->rank] = c->base;
    ctx->comm = c;
    return LQCD_OK;
}

int comm_destroy(lqcd_ctx *ctx) {
    CommState *c = ctx->comm;
    if (!c) return LQCD_OK;
    for (int r = 0; r < ctx->nranks; r++)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
    for (int r = 0; r < ctx->nranks; r++)
        if (c->gauge_opened[r]) cudaIpcCloseMemHandle((void *)c->gauge_peer[r]);
    cudaFree(c->base);
    cudaFree(c->bsites);
    cudaFree(c->cta_order);
    delete c;
    ctx->comm = nullptr;
    return LQCD_OK;
}

extern "C" int lqcd_comm_export(lqcd_ctx *ctx, void *handle_out) {
    if (!ctx || !handle_out) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LQCD_TRY(comm_alloc(ctx));
    HandleBlob b;
    memset(&b, 0, sizeof b);
    b.magic = LQCD_HANDLE_MAGIC; b.rank = ctx->rank; b.pid = (int64_t)getpid(); b.device = ctx->device;
    b.raw_ptr = (uint64_t)(uintptr_t)ctx->comm->base; b.bytes = ctx->comm->bytes;
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&b.ipc, ctx->comm->base));
    // optional second mapping (never fatal: only the multi-rank clover build needs it)
    if (cudaIpcGetMemHandle(&b.gauge_ipc, ctx->gauge) == cudaSuccess) { b.has_gauge = 1; b.gauge_raw = (uint64_t)(uintptr_t)ctx->gauge; }
    else cudaGetLastError();
    memset(handle_out, 0, LQCD_IPC_HANDLE_BYTES);
    memcpy(handle_out, &b, sizeof b);
    return LQCD_OK;
}

extern "C" int lqcd_comm_connect(lqcd_ctx *ctx, const void *all_handles) {
    if (!ctx || !all_handles) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (ctx->nranks == 1) return LQCD_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LQCD_TRY(comm_alloc(ctx));
    CommState *c = ctx->comm;
    for (int r = 0; r < ctx->nranks; r++) {
        HandleBlob b;
        memcpy(&b, (const char *)all_handles + (size_t)r * LQCD_IPC_HANDLE_BYTES, sizeof b);
        if (b.magic != LQCD_HANDLE_MAGIC || b.rank != r) return lqcd_fail(ctx, LQCD_ERR_COMM, "handle %d is malformed (rank order?)", r);
        if (b.bytes != c->bytes) return lqcd_fail(ctx, LQCD_ERR_COMM, "rank %d has a different comm layout (%llu vs %zu bytes)", r, (unsigned long long)b.bytes, c->bytes);
        if (r == ctx->rank) continue;
        if (c->peer[r]) continue;
        if (b.pid == (int64_t)getpid()) {            // same process (several contexts in one process): raw pointer
            if (b.device != ctx->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return lqcd_fail(ctx, LQCD_ERR_COMM, "cudaDeviceEnablePeerAccess(%d) -> %s", b.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            c->peer[r] = (char *)(uintptr_t)b.raw_ptr;
        } else {
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return lqcd_fail(ctx, LQCD_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) -> %s (NVLink/P2P peer access is required)", r, cudaGetErrorString(e));
            c->peer[r] = (char *)p; c->opened[r] = true;
        }
        // link array of rank r (optional; failures leave gauge_peer[r] null and only disable the multi-rank clover build)
        if (b.has_gauge && !c->gauge_peer[r]) {
            if (b.pid == (int64_t)getpid()) c->gauge_peer[r] = (const cplx *)(uintptr_t)b.gauge_raw;
            else {
                void *p = nullptr;
                if (cudaIpcOpenMemHandle(&p, b.gauge_ipc, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) { c->gauge_peer[r] = (const cplx *)p; c->gauge_opened[r] = true; }
                else cudaGetLastError();
            }
        }
    }
    c->gauge_peer[ctx->rank] = ctx->gauge;
    CommRed &cr = ctx->red.cr;
    cr.nranks = ctx->nranks; cr.rank = ctx->rank;
    cr.seq = (unsigned long long *)(c->base + c->off_seq);
    cr.err = (int *)(c->base + c->off_err);
    {
        double secs = LQCD_SPIN_TIMEOUT_DEFAULT_S;
        if (const char *e = getenv("LQCD_COMM_TIMEOUT_S")) { double v = atof(e); if (v > 0.01 && v < 3600.0) secs = v; }
        cr.timeout_cycles = (long long)(secs * 1.9e9);
    }
    for (int r = 0; r < ctx->nranks; r++) {
        cr.vals[r] = (double *)(c->peer[r] + c->off_red_vals);
        cr.flags[r] = (unsigned long long *)(c->peer[r] + c->off_red_flags);
    }
    c->connected = true;
    return LQCD_OK;
}

// every rank's link array as seen from this GPU.  The CALLER must have barriered the ranks after their gauge uploads.
int comm_link_view(lqcd_ctx *ctx, const cplx **bases) {
    CommState *c = ctx->comm;
    if (!c || !c->connected) return lqcd_fail(ctx, LQCD_ERR_COMM, "multi-rank context is not connected (lqcd_comm_export / lqcd_comm_connect)");
    for (int r = 0; r < ctx->nranks; r++) {
        if (!c->gauge_peer[r]) return lqcd_fail(ctx, LQCD_ERR_COMM, "link array of rank %d is not peer mapped (clover build across ranks needs it)", r);
        bases[r] = c->gauge_peer[r];
    }
    return LQCD_OK;
}

// fermion-force halo descriptors for force call number ++force_seq (see ForceHalo in lqcd_internal.cuh / force.cu)
int comm_force_halo(lqcd_ctx *ctx, ForceHalo *F) {
    CommState *c = ctx->comm;
    if (!c || !c->connected) return lqcd_fail(ctx, LQCD_ERR_COMM, "multi-rank context is not connected (lqcd_comm_export / lqcd_comm_connect)");
    memset(F, 0, sizeof *F);
    F->seq = ++c->force_seq;
    F->ticket = (unsigned int *)(c->base + c->off_ticket);
    F->err = (int *)(c->base + c->off_err);
    F->timeout_cycles = ctx->red.cr.timeout_cycles;
    int n = 0;
    for (int mu = 0; mu < 4; mu++) {
        F->start[mu] = n;
        F->plast[mu] = ctx->pcoord[mu] == ctx->procgrid[mu] - 1;
        if (!ctx->g.part[mu]) continue;
        n += c->face[mu];
        const int lo = c->nbr[mu][0];
        F->send[mu] = (cplx *)(c->peer[lo] + c->force_off[mu]);
        F->send_flag[mu] = (unsigned long long *)(c->peer[lo] + c->off_force_flags) + mu;
        F->recv[mu] = (const cplx *)(c->base + c->force_off[mu]);
        F->recv_flag[mu] = (const unsigned long long *)(c->base + c->off_force_flags) + mu;
    }
    F->start[4] = n;
    return LQCD_OK;
}

int comm_allreduce_sum(lqcd_ctx *, double *, int) { return LQCD_OK; }   // reductions are already global (in-kernel)

int comm_check_error(lqcd_ctx *ctx) {
    if (ctx->nranks == 1 || !ctx->comm) return LQCD_OK;
    int err = 0;
    CUDA_TRY(ctx, cudaMemcpy(&err, ctx->comm->base + ctx->comm->off_err, sizeof err, cudaMemcpyDeviceToHost));
    if (err) return lqcd_fail(ctx, LQCD_ERR_COMM, "peer wait timed out on the device (code %d: 1xxxxxx = halo flag [dir*2+side][seq], 2xxxxxx = all-reduce [seq], 3xxxxxx = force halo [dir][seq]; host halo_seq = %llu)", err, ctx->comm->halo_seq);
    return LQCD_OK;
}

// ---- kernels -------------------------------------------------------------------------------------------
struct HaloArgs {
    const cplx *in;            // pack: source spinor.  exterior: unused
    cplx *out;                 // exterior: y (read-modify-write)
    const cplx *gauge;
    Geom g;
    int kind, dagger;
    double coef;               // Wilson: -kappa ; staggered: sign
    double bc[4];
    int pfirst[4], plast[4];   // this rank sits on the global low / high boundary in mu
    HaloOut hout;              // pack: destinations, flags, CTA prefix, ticket, seq
    const cplx *recv[4][2];    // exterior: [mu][0] data from lower nbr (for my low face), [1] from upper nbr (for my high face)
    const unsigned long long *recv_flag[4][2];
    unsigned long long seq;
    int *err;
    const SolverState *st;
    int use_state;
    const int *bsites;         // exterior: table of boundary sites (on at least one partitioned face)
    int nbsites;
    DslashFuse fuse;           // exterior: fused epilogue (reductions finished here)
    Reduce red;
    unsigned int part_offset;  // exterior: partials deposited by the interior kernel precede ours
};

__global__ void __launch_bounds__(128) halo_pack_kernel(const HaloArgs A) {
    if (A.use_state && A.st->done) return;
    halo_pack_cta(A.g, A.kind, A.dagger, A.in, A.gauge, A.hout, blockIdx.x);
}

// Wilson: add the off-rank hop(s) of direction MU at face site s into acc (12 complex, in units of "hopping sum").
template <int MU>
__device__ __forceinline__ void wilson_ext_dir(const HaloArgs &A, cplx (&acc)[12], int s, int side, int f) {
    const cplx *src = A.recv[MU][side] + (size_t)(f >> 5) * (6 * 32) + (f & 31);
    cplx h0[3], h1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { h0[c] = __ldcg(src + c * 32); h1[c] = __ldcg(src + (3 + c) * 32); }
    const double phase = side == 0 ? (A.pfirst[MU] ? A.bc[MU] : 1.0) : (A.plast[MU] ? A.bc[MU] : 1.0);
    if (phase != 1.0) {
#pragma unroll
        for (int c = 0; c < 3; c++) { h0[c] = cscale(phase, h0[c]); h1[c] = cscale(phase, h1[c]); }
    }
    if (side == 0) {                        // my low face, backward hop: data is already U^dag P psi
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (A.dagger) reconstruct<MU, -1>(acc, a, h0[a], h1[a]);
            else          reconstruct<MU, +1>(acc, a, h0[a], h1[a]);
        }
    } else {                                // my high face, forward hop: apply U_mu(n)
        const cplx *lk = A.gauge + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cplx g0 = cmake(0, 0), g1 = cmake(0, 0);
#pragma unroll
            for (int b = 0; b < 3; b++) {
                cplx u = ldg128(lk + (a * 3 + b) * 32);
                cfma(g0, u, h0[b]); cfma(g1, u, h1[b]);
            }
            if (A.dagger) reconstruct<MU, +1>(acc, a, g0, g1);
            else          reconstruct<MU, -1>(acc, a, g0, g1);
        }
    }
}

template <int MU>
__device__ __forceinline__ void stag_ext_dir(const HaloArgs &A, cplx (&acc)[3], int s, int side, int f, double eta) {
    const cplx *src = A.recv[MU][side] + (size_t)(f >> 5) * (6 * 32) + (f & 31);
    cplx h[3];
#pragma unroll
    for (int c = 0; c < 3; c++) h[c] = __ldcg(src + c * 32);
    if (side == 0) {
        const double cf = -0.5 * eta * (A.pfirst[MU] ? A.bc[MU] : 1.0);
#pragma unroll
        for (int c = 0; c < 3; c++) { acc[c].x = fma(cf, h[c].x, acc[c].x); acc[c].y = fma(cf, h[c].y, acc[c].y); }
    } else {
        const double cf = 0.5 * eta * (A.plast[MU] ? A.bc[MU] : 1.0);
        const cplx *lk = A.gauge + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cplx g = cmake(0, 0);
#pragma unroll
            for (int b = 0; b < 3; b++) cfma(g, ldg128(lk + (a * 3 + b) * 32), h[b]);
            acc[a].x = fma(cf, g.x, acc[a].x); acc[a].y = fma(cf, g.y, acc[a].y);
        }
    }
}

#define EXT_DIR_W(MU, coord, dim)                                                            \
    if (A.g.part[MU]) {                                                                      \
        if ((coord) == 0) wilson_ext_dir<MU>(A, acc, s, 0, face_index<MU>(A.g, x, y, z, t)); \
        if ((coord) == (dim)-1) wilson_ext_dir<MU>(A, acc, s, 1, face_index<MU>(A.g, x, y, z, t)); \
    }
#define EXT_DIR_S(MU, coord, dim, eta)                                                       \
    if (A.g.part[MU]) {                                                                      \
        if ((coord) == 0) stag_ext_dir<MU>(A, acc, s, 0, face_index<MU>(A.g, x, y, z, t), eta); \
        if ((coord) == (dim)-1) stag_ext_dir<MU>(A, acc, s, 1, face_index<MU>(A.g, x, y, z, t), eta); \
    }

// One thread per BOUNDARY site (precomputed table): all its off-rank hops are added in registers and y is
// updated once, so sites on several faces (edges/corners of a T x Z decomposition) have no write race, and the
// fused <w,y>, |y|^2 reductions of these sites are finished here together with the interior kernel's partials.
__global__ void __launch_bounds__(128) halo_exterior_kernel(const HaloArgs A) {
    if (A.use_state && A.st->done) return;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        int good = 1;
        for (int m = 0; m < 4 && good; m++) {
            if (!A.g.part[m]) continue;
            for (int side = 0; side < 2 && good; side++)
                while (ld_acquire_sys(A.recv_flag[m][side]) < A.seq)
                    if (clock64() - t0 > A.red.cr.timeout_cycles) { good = 0; *A.err = 1; break; }
        }
        ok = good;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[3] = {0.0, 0.0, 0.0};
    if (ok && i < A.nbsites) {
        const int s = A.bsites[i];
        int r = s;
        const int x = r % A.g.X; r /= A.g.X;
        const int y = r % A.g.Y; r /= A.g.Y;
        const int z = r % A.g.Z;
        const int t = r / A.g.Z;
        if (A.kind == LQCD_WILSON) {
            cplx acc[12];
#pragma unroll
            for (int k = 0; k < 12; k++) acc[k] = cmake(0, 0);
            EXT_DIR_W(0, x, A.g.X) EXT_DIR_W(1, y, A.g.Y) EXT_DIR_W(2, z, A.g.Z) EXT_DIR_W(3, t, A.g.T)
            const size_t base = (size_t)(s >> 5) * (12 * 32) + (s & 31);
#pragma unroll
            for (int k = 0; k < 12; k++) {
                cplx v = A.out[base + k * 32];
                v.x = fma(A.coef, acc[k].x, v.x); v.y = fma(A.coef, acc[k].y, v.y);
                A.out[base + k * 32] = v;
                if (A.fuse.dot_with) {
                    cplx w = ldg128(A.fuse.dot_with + base + k * 32);
                    red[0] = fma(w.x, v.x, red[0]); red[0] = fma(w.y, v.y, red[0]);
                    red[1] = fma(w.x, v.y, red[1]); red[1] = fma(-w.y, v.x, red[1]);
                }
                red[2] = fma(v.x, v.x, red[2]); red[2] = fma(v.y, v.y, red[2]);
            }
        } else {
            cplx acc[3];
#pragma unroll
            for (int k = 0; k < 3; k++) acc[k] = cmake(0, 0);
            const int gx = x + A.g.o[0], gy = y + A.g.o[1], gz = z + A.g.o[2];
            EXT_DIR_S(0, x, A.g.X, 1.0)
            EXT_DIR_S(1, y, A.g.Y, ((gx & 1) ? -1.0 : 1.0))
            EXT_DIR_S(2, z, A.g.Z, (((gx + gy) & 1) ? -1.0 : 1.0))
            EXT_DIR_S(3, t, A.g.T, (((gx + gy + gz) & 1) ? -1.0 : 1.0))
            const size_t base = (size_t)(s >> 5) * (3 * 32) + (s & 31);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                cplx v = A.out[base + k * 32];
                v.x = fma(A.coef, acc[k].x, v.x); v.y = fma(A.coef, acc[k].y, v.y);
                A.out[base + k * 32] = v;
                if (A.fuse.dot_with) {
                    cplx w = ldg128(A.fuse.dot_with + base + k * 32);
                    red[0] = fma(w.x, v.x, red[0]); red[0] = fma(w.y, v.y, red[0]);
                    red[1] = fma(w.x, v.y, red[1]); red[1] = fma(-w.y, v.x, red[1]);
                }
                red[2] = fma(v.x, v.x, red[2]); red[2] = fma(v.y, v.y, red[2]);
            }
        }
    }
    if (A.fuse.dot_with || A.fuse.want_norm)
        grid_reduce_finish<3>(red, A.red, A.fuse.finish, A.part_offset, A.part_offset + gridDim.x, 1);
}

int comm_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse) {
    CommState *c = ctx->comm;
    if (!c || !c->connected) return lqcd_fail(ctx, LQCD_ERR_COMM, "multi-rank context is not connected (lqcd_comm_export / lqcd_comm_connect)");
    const Geom &g = ctx->g;
    HaloArgs A;
    memset(&A, 0, sizeof A);
    A.in = x; A.out = y; A.gauge = ctx->gauge; A.g = g; A.kind = op->kind; A.dagger = dagger;
    A.coef = (op->kind == LQCD_WILSON) ? -op->kappa : (dagger ? -1.0 : 1.0);
    const unsigned long long seq = ++c->halo_seq;
    const int slot = (int)(seq & 1);
    A.seq = seq;
    A.hout.seq = seq;
    {
        static int gf = -1;
        if (gf < 0) { const char *e = getenv("LQCD_PACK_FENCE"); gf = (e && e[0] == 'g') ? 1 : 0; }
        A.hout.gpu_fence = gf;
    }
    A.hout.ticket = (unsigned int *)(c->base + c->off_ticket);
    A.err = (int *)(c->base + c->off_err);
    A.st = ctx->red.st; A.use_state = fuse ? fuse->use_state : 0;
    A.bsites = c->bsites; A.nbsites = c->nbsites;
    A.red = ctx->red;
    int ncta = 0;
    for (int mu = 0; mu < 4; mu++) {
        A.bc[mu] = op->bc[mu];
        A.pfirst[mu] = ctx->pcoord[mu] == 0; A.plast[mu] = ctx->pcoord[mu] == ctx->procgrid[mu] - 1;
        A.hout.cta0[mu] = ncta;      // number of CTAs of partitioned directions < mu (non-partitioned ones own zero CTAs)
        if (!g.part[mu]) continue;
        ncta += (2 * c->face[mu] + 127) / 128;
        const int lo = c->nbr[mu][0], hi = c->nbr[mu][1];
        // my low face feeds the LOWER neighbour's "from upper" (side 1) slot; my high face the UPPER neighbour's side 0
        A.hout.send[mu][0] = (cplx *)(c->peer[lo] + c->halo_off[mu][1][slot]);
        A.hout.send[mu][1] = (cplx *)(c->peer[hi] + c->halo_off[mu][0][slot]);
        A.hout.send_flag[mu][0] = (unsigned long long *)(c->peer[lo] + c->off_halo_flags) + (mu * 2 + 1) * 2 + slot;
        A.hout.send_flag[mu][1] = (unsigned long long *)(c->peer[hi] + c->off_halo_flags) + (mu * 2 + 0) * 2 + slot;
        for (int side = 0; side < 2; side++) {
            A.recv[mu][side] = (const cplx *)(c->base + c->halo_off[mu][side][slot]);
            A.recv_flag[mu][side] = (const unsigned long long *)(c->base + c->off_halo_flags) + (mu * 2 + side) * 2 + slot;
        }
    }
    A.hout.cta0[4] = ncta;
    // Default: SELF-PACKING Dslash kernel -- the first npack CTAs of the kernel itself ship this application's halo
    // (halo_pack.cuh), the interior tiles follow, the face tiles (last) consume the neighbours' slots: one launch,
    // no second stream, no events.  LQCD_SELF_PACK=0 falls back to a separate pack kernel on the priority stream.
    // Measured (2xB200): local volume 32.32.16.8 -> 39.0 us self-packing vs 41.7 us separate pack; local volume
    // 32.32.32.16 -> 122.5-126.5 vs 119.9 us (the leading pack CTAs delay the first interior wave).  8xB200, grid 1.1.2.4 (two
    // partitioned directions, local volume 32.32.16.8): 52.0 us / 5257 CG it/s self-packing vs 51.1 us / 5658 it/s separate pack
    // (profiles/r1g_scale_n8{,_sep}.json; identical iterates) -- with two face pairs the leading pack phase is twice as long.
    // Default: self-pack for local volumes up to 2^18 sites when ONE direction is partitioned (and for tiny local volumes below
    // 2^15 sites, where the difference is noise and the hardware-verified test configurations stay exactly as verified),
    // separate pack kernel otherwise; LQCD_SELF_PACK=0/1 forces either.
    static int relaxed_poll = -1;
    if (relaxed_poll < 0) { const char *e = getenv("LQCD_HALO_POLL"); relaxed_poll = (e && e[0] == 'r') ? 1 : 0; }
    static int self_pack_env = -2;
    if (self_pack_env == -2) { const char *e = getenv("LQCD_SELF_PACK"); self_pack_env = e ? (atoi(e) != 0) : -1; }
    const int npart = g.part[0] + g.part[1] + g.part[2] + g.part[3];
    const int self_pack = self_pack_env >= 0 ? self_pack_env : (g.V <= (1 << 18) && (npart == 1 || g.V < (1 << 15)));
    if (self_pack) {
        const int bs = 32 * g.wpc;
        HaloOut O = A.hout;
        int np = 0;
        for (int mu = 0; mu < 4; mu++) { O.cta0[mu] = np; if (g.part[mu]) np += (2 * c->face[mu] + bs - 1) / bs; }
        O.cta0[4] = np;
        HaloIn H;
        memset(&H, 0, sizeof H);
        for (int mu = 0; mu < 4; mu++) {
            H.pfirst[mu] = A.pfirst[mu]; H.plast[mu] = A.plast[mu];
            for (int side = 0; side < 2; side++) { H.recv[mu][side] = A.recv[mu][side]; H.recv_flag[mu][side] = A.recv_flag[mu][side]; }
        }
        H.seq = seq; H.err = A.err; H.cta_order = c->cta_order; H.n_interior = c->n_interior;
        H.timeout_cycles = ctx->red.cr.timeout_cycles;
        H.relaxed_poll = relaxed_poll;
        if (op->kind == LQCD_WILSON) return launch_wilson_dslash(ctx, op, y, x, dagger, fuse, ctx->stream, &H, &O);
        return launch_staggered_dslash(ctx, op, y, x, dagger, fuse, ctx->stream, &H, &O);
    }
    // pack on the second stream (overlaps the interior kernel); it needs x, which earlier main-stream work produced
    static int two_streams = -1, timing = -1;
    static cudaEvent_t te[4];
    static double tacc[3] = {0, 0, 0};
    static long tcount = 0;
    if (two_streams < 0) { const char *e = getenv("LQCD_PACK_STREAM"); two_streams = (e && atoi(e) == 0) ? 0 : 1; }
    // LQCD_COMM_TIMING=2: non-intrusive timeline (events recorded on both streams, no host sync until 200 samples)
    static const int TLN = 200;
    static cudaEvent_t tl[4][TLN];      // 0 pack start, 1 pack end, 2 dslash start, 3 dslash end
    static int tln = 0, timeline = 0;
    if (timing < 0) {
        const char *e = getenv("LQCD_COMM_TIMING"); timing = (e && atoi(e) == 1) ? 1 : 0;
        timeline = (e && atoi(e) == 2) ? 1 : 0;
        if (timing) { two_streams = 0; for (int i = 0; i < 4; i++) cudaEventCreate(&te[i]); }
        if (timeline) for (int i = 0; i < 4; i++) for (int j = 0; j < TLN; j++) cudaEventCreate(&tl[i][j]);
    }
    const bool tlrec = timeline && tln < TLN;
    if (timing) cudaEventRecord(te[0], ctx->stream);
    cudaStream_t ps = two_streams ? ctx->stream2 : ctx->stream;
    if (two_streams) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_int, ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ps, ctx->ev_int, 0));
    }
    if (tlrec) cudaEventRecord(tl[0][tln], ps);
    halo_pack_kernel<<<ncta, 128, 0, ps>>>(A);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    if (tlrec) cudaEventRecord(tl[1][tln], ps);
    if (two_streams) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_pack, ps));
    if (timing) cudaEventRecord(te[1], ctx->stream);
    static int fused = -1;
    if (fused < 0) { const char *e = getenv("LQCD_FUSED_HALO"); fused = (e && atoi(e) == 0) ? 0 : 1; }
    if (fused) {
        // ONE kernel on the critical path: the Dslash kernel itself consumes the halo slots (face tiles run last and
        // wait on the neighbours' flags in-kernel); reductions finish there as on a single GPU.
        HaloIn H;
        memset(&H, 0, sizeof H);
        for (int mu = 0; mu < 4; mu++) {
            H.pfirst[mu] = A.pfirst[mu]; H.plast[mu] = A.plast[mu];
            for (int side = 0; side < 2; side++) { H.recv[mu][side] = A.recv[mu][side]; H.recv_flag[mu][side] = A.recv_flag[mu][side]; }
        }
        H.seq = seq; H.err = A.err; H.cta_order = c->cta_order; H.n_interior = c->n_interior;
        H.timeout_cycles = ctx->red.cr.timeout_cycles;
        H.relaxed_poll = relaxed_poll;
        if (tlrec) cudaEventRecord(tl[2][tln], ctx->stream);
        if (op->kind == LQCD_WILSON) LQCD_TRY(launch_wilson_dslash(ctx, op, y, x, dagger, fuse, ctx->stream, &H));
        else                         LQCD_TRY(launch_staggered_dslash(ctx, op, y, x, dagger, fuse, ctx->stream, &H));
        if (tlrec) {
            cudaEventRecord(tl[3][tln], ctx->stream);
            if (++tln == TLN) {
                cudaEventSynchronize(tl[3][TLN - 1]); cudaEventSynchronize(tl[1][TLN - 1]);
                double a[6] = {0, 0, 0, 0, 0, 0};
                int n = 0;
                for (int k = TLN / 2; k < TLN; k++, n++) {          // second half: steady state
                    float ms;
                    cudaEventElapsedTime(&ms, tl[0][k], tl[1][k]); a[0] += ms;          // pack duration
                    cudaEventElapsedTime(&ms, tl[2][k], tl[3][k]); a[1] += ms;          // dslash duration
                    cudaEventElapsedTime(&ms, tl[3][k - 1], tl[2][k]); a[2] += ms;      // gap: previous dslash end -> this start
                    cudaEventElapsedTime(&ms, tl[2][k], tl[0][k]); a[3] += ms;          // pack start relative to dslash start
                    cudaEventElapsedTime(&ms, tl[2][k], tl[1][k]); a[4] += ms;          // pack end relative to dslash start
                    cudaEventElapsedTime(&ms, tl[3][k - 1], tl[3][k]); a[5] += ms;      // period
                }
                fprintf(stderr, "[lqcd timeline rank %d] pack %.1f us | dslash %.1f us | gap %.1f us | pack start %+.1f us, end %+.1f us after dslash start | period %.1f us\n",
                        ctx->rank, 1e3 * a[0] / n, 1e3 * a[1] / n, 1e3 * a[2] / n, 1e3 * a[3] / n, 1e3 * a[4] / n, 1e3 * a[5] / n);
            }
        }
        if (two_streams) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pack, 0));
        if (timing) {
            cudaEventRecord(te[2], ctx->stream);
            cudaEventSynchronize(te[2]);
            for (int i = 0; i < 2; i++) { float ms; cudaEventElapsedTime(&ms, te[i], te[i + 1]); tacc[i] += ms; }
            if (++tcount % 64 == 0)
                fprintf(stderr, "[lqcd comm timing rank %d] pack %.1f us  fused dslash %.1f us (mean of %ld)\n",
                        ctx->rank, 1e3 * tacc[0] / tcount, 1e3 * tacc[1] / tcount, tcount);
        }
        return LQCD_OK;
    }
    if (fuse && fuse->axpy_r)
        return lqcd_fail(ctx, LQCD_ERR_ARG, "LQCD_FUSED_HALO=0 (separate exterior kernel) needs LQCD_CG_FUSE=0 as well");
    // interior: all sites with off-rank hops masked; reductions over non-face sites deposited as partials
    const bool want_red = fuse && (fuse->dot_with || fuse->want_norm);
    DslashFuse f2 = DslashFuse();
    if (fuse) f2 = *fuse;
    f2.interior_only = 1;
    if (op->kind == LQCD_WILSON) LQCD_TRY(launch_wilson_dslash(ctx, op, y, x, dagger, &f2, ctx->stream));
    else                         LQCD_TRY(launch_staggered_dslash(ctx, op, y, x, dagger, &f2, ctx->stream));
    if (timing) cudaEventRecord(te[2], ctx->stream);
    // x must not be overwritten by later main-stream kernels before the pack has read it
    if (two_streams) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pack, 0));
    A.fuse = f2; A.fuse.interior_only = 0;
    A.part_offset = want_red ? (unsigned int)((g.nblk + g.wpc - 1) / g.wpc) : 0u;
    halo_exterior_kernel<<<(c->nbsites + 127) / 128, 128, 0, ctx->stream>>>(A);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    if (timing) {       // development aid: serialises the host; prints the mean split every 64 applications
        cudaEventRecord(te[3], ctx->stream);
        cudaEventSynchronize(te[3]);
        for (int i = 0; i < 3; i++) { float ms; cudaEventElapsedTime(&ms, te[i], te[i + 1]); tacc[i] += ms; }
        if (++tcount % 64 == 0)
            fprintf(stderr, "[lqcd comm timing rank %d] pack %.1f us  interior %.1f us  exterior %.1f us (mean of %ld)\n",
                    ctx->rank, 1e3 * tacc[0] / tcount, 1e3 * tacc[1] / tcount, 1e3 * tacc[2] / tcount, tcount);
    }
    return LQCD_OK;
}

// ---- pure geometry helper (no GPU needed): used by the host-side tests of the decomposition ----------------
extern "C" int lqcd_decompose(const int gd[4], const int pg[4], int rank, int local_dims[4], int origin[4], int nbr_lo[4], int nbr_hi[4]) {
    if (!gd || !pg || !local_dims || !origin || !nbr_lo || !nbr_hi) return LQCD_ERR_ARG;
    int n = pg[0] * pg[1] * pg[2] * pg[3];
    if (n < 1 || rank < 0 || rank >= n) return LQCD_ERR_ARG;
    int pc[4], r = rank;
    for (int i = 0; i < 4; i++) {
        if (pg[i] < 1 || gd[i] % pg[i] != 0) return LQCD_ERR_ARG;
        pc[i] = r % pg[i]; r /= pg[i];
        local_dims[i] = gd[i] / pg[i]; origin[i] = pc[i] * local_dims[i];
    }
    for (int mu = 0; mu < 4; mu++) {
        int q[4] = {pc[0], pc[1], pc[2], pc[3]};
        q[mu] = (pc[mu] + pg[mu] - 1) % pg[mu]; nbr_lo[mu] = q[0] + pg[0] * (q[1] + pg[1] * (q[2] + pg[2] * q[3]));
        q[mu] = (pc[mu] + 1) % pg[mu];          nbr_hi[mu] = q[0] + pg[0] * (q[1] + pg[1] * (q[2] + pg[2] * q[3]));
    }
    return LQCD_OK;
}
