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

#define LQCD_TIMING_SLOTS 1024

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
    unsigned long long *timing;       // LQCD_COMM_TIMING=1: [LQCD_TIMING_SLOTS][8] phase stamps (see comm_timing_report)
    unsigned long long timing_first;  // halo_seq of slot 0
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
        if (e != cudaSuccess) { cudaFree(c->base); delete c; return lqcd_fail(ctx, LQCD_ERR_CUDA, "cta order table -> %s", cudaGetErrorString(e)); }
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
    cudaFree(c->timing);
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
    const cplx *in;            // source spinor
    const cplx *gauge;
    Geom g;
    int kind, dagger;
    HaloOut hout;              // destinations, flags, CTA prefix, ticket, seq
    const SolverState *st;
    int use_state;
};

__global__ void __launch_bounds__(128) halo_pack_kernel(const HaloArgs A) {
    if (A.use_state && A.st->done) return;
    halo_pack_cta(A.g, A.kind, A.dagger, A.in, A.gauge, A.hout, blockIdx.x);
}

bool wilson_tmarch_ok(const lqcd_ctx *ctx, const lqcd_op *op);      // wilson_tmarch.cu

// One operator application across ranks.  The halo producer (halo_pack.cuh) either leads the Dslash kernel itself ("self-pack":
// the first npack CTAs of the grid) or runs as its own kernel on the highest-priority stream; the Dslash kernel consumes the
// neighbours' slots in its face tiles, which are last in the CTA order and wait on the sequence flags.  (Round 1 also had a
// three-kernel form with a separate exterior kernel and host-timed variants of it: slower at every N, removed in round 2.)
int comm_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse) {
    CommState *c = ctx->comm;
    if (!c || !c->connected) return lqcd_fail(ctx, LQCD_ERR_COMM, "multi-rank context is not connected (lqcd_comm_export / lqcd_comm_connect)");
    const Geom &g = ctx->g;
    const unsigned long long seq = ++c->halo_seq;
    const int slot = (int)(seq & 1);
    // measured on 2 / 8 B200 (profiles/r1g_*, round-1 experiments leg): a gpu-scope fence per pack CTA with ONE cumulative system
    // fence by the last pack CTA is faster than a system fence per CTA at every N (N = 2: 0.115 vs 0.125 ms, N = 8: 0.042 vs 0.045 ms)
    static int sys_fence = -1, timing = -1, self_pack_env = -2, spt_env = -1, tune = -1;
    if (tune < 0) { const char *e = getenv("LQCD_COMM_TUNE"); tune = (e && atoi(e) == 1) ? 1 : 0; }
    if (tune) { timing = -1; self_pack_env = -2; spt_env = -1; }      // tools/comm_tune.py changes the knobs between timed runs of one process
    if (sys_fence < 0) { const char *e = getenv("LQCD_PACK_FENCE"); sys_fence = (e && e[0] == 's') ? 1 : 0; }
    if (timing < 0) { const char *e = getenv("LQCD_COMM_TIMING"); timing = (e && atoi(e) >= 1) ? 1 : 0; }
    if (self_pack_env == -2) { const char *e = getenv("LQCD_SELF_PACK"); self_pack_env = e ? (atoi(e) != 0) : -1; }
    if (spt_env < 0) { const char *e = getenv("LQCD_PACK_SPT"); spt_env = (e && atoi(e) >= 1 && atoi(e) <= 64) ? atoi(e) : 0; }
    unsigned long long *tm = nullptr;
    if (timing) {
        if (!c->timing) {
            CUDA_TRY(ctx, cudaMalloc(&c->timing, sizeof(unsigned long long) * 8 * LQCD_TIMING_SLOTS));
            std::vector<unsigned long long> init(8 * LQCD_TIMING_SLOTS, 0ull);
            for (int i = 0; i < LQCD_TIMING_SLOTS; i++) { init[i * 8 + 0] = init[i * 8 + 2] = init[i * 8 + 4] = ~0ull; }
            CUDA_TRY(ctx, cudaMemcpy(c->timing, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
            c->timing_first = seq;
        }
        if (seq - c->timing_first < LQCD_TIMING_SLOTS) tm = c->timing + (seq - c->timing_first) * 8;
    }
    const bool tmarch = op->kind == LQCD_WILSON && wilson_tmarch_ok(ctx, op);     // consumes the slots itself, fed by the separate pack kernel
    // Pack placement, measured on one 8 x B200 box at 32^4 (round 2, profiles/r2d_comm_tune_sweep.txt, us per Wilson application):
    //   local volume 2^17 (8 GPUs): separate pack kernel 40.6 | pack CTAs leading the Dslash kernel, 1 / 2 / 4 / 8 / 16 face sites per
    //                               pack thread 38-58 / 36.8 / 33.4 / 35.3 / 44.6
    //   local volume 2^18 (4 GPUs): separate 66.3 (4 sites per thread 62.8) | leading 70.2 / 66.0 / 64.8 / 63.9 / 65.2
    //   local volume 2^19 (2 GPUs): separate 108.4 (4 sites per thread 111.7) | leading 120.7 / 115.6 / 114.5 / 115.2 / 115.4
    // (staggered: leading with 2-4 sites per thread wins up to 2^18, separate with 4 at 2^19).  The in-kernel timeline shows why: a
    // separate pack kernel only gets SM slots as first-wave Dslash CTAs retire (its CTAs start 6-10 us late), which is hidden behind
    // a long interior phase but not behind the 15-19 us of the 8-GPU local volume; leading pack CTAs start at once, and fewer, longer
    // ones leave the rest of the first wave to the interior tiles.
    const int self_limit = op->kind == LQCD_WILSON ? (1 << 17) : (1 << 18);
    const int self_pack = tmarch ? 0 : (self_pack_env >= 0 ? self_pack_env : (g.V <= self_limit));
    const int pbs = self_pack ? 32 * g.wpc : 128;                                   // threads per pack CTA
    const int spt = spt_env ? spt_env : (self_pack ? 4 : (g.V <= (1 << 18) || op->kind == LQCD_STAGGERED ? 4 : 1));   // face sites per pack thread
    HaloOut O;
    HaloIn H;
    memset(&O, 0, sizeof O);
    memset(&H, 0, sizeof H);
    O.seq = seq; O.gpu_fence = !sys_fence; O.ticket = (unsigned int *)(c->base + c->off_ticket); O.timing = tm; O.spt = spt;
    H.seq = seq; H.err = (int *)(c->base + c->off_err); H.cta_order = c->cta_order; H.n_interior = c->n_interior;
    H.timeout_cycles = ctx->red.cr.timeout_cycles; H.timing = tm;
    int ncta = 0;
    for (int mu = 0; mu < 4; mu++) {
        H.pfirst[mu] = ctx->pcoord[mu] == 0; H.plast[mu] = ctx->pcoord[mu] == ctx->procgrid[mu] - 1;
        O.cta0[mu] = ncta;           // number of pack CTAs of partitioned directions < mu (non-partitioned ones own zero CTAs)
        if (!g.part[mu]) continue;
        ncta += (2 * c->face[mu] + pbs * spt - 1) / (pbs * spt);
        const int lo = c->nbr[mu][0], hi = c->nbr[mu][1];
        // my low face feeds the LOWER neighbour's "from upper" (side 1) slot; my high face the UPPER neighbour's side 0
        O.send[mu][0] = (cplx *)(c->peer[lo] + c->halo_off[mu][1][slot]);
        O.send[mu][1] = (cplx *)(c->peer[hi] + c->halo_off[mu][0][slot]);
        O.send_flag[mu][0] = (unsigned long long *)(c->peer[lo] + c->off_halo_flags) + (mu * 2 + 1) * 2 + slot;
        O.send_flag[mu][1] = (unsigned long long *)(c->peer[hi] + c->off_halo_flags) + (mu * 2 + 0) * 2 + slot;
        for (int side = 0; side < 2; side++) {
            H.recv[mu][side] = (const cplx *)(c->base + c->halo_off[mu][side][slot]);
            H.recv_flag[mu][side] = (const unsigned long long *)(c->base + c->off_halo_flags) + (mu * 2 + side) * 2 + slot;
        }
    }
    O.cta0[4] = ncta;
    if (self_pack) {
        if (op->kind == LQCD_WILSON) return launch_wilson_dslash(ctx, op, y, x, dagger, fuse, ctx->stream, &H, &O);
        return launch_staggered_dslash(ctx, op, y, x, dagger, fuse, ctx->stream, &H, &O);
    }
    // pack on the priority stream (overlaps the interior tiles); it needs x, which earlier main-stream work produced
    HaloArgs A;
    memset(&A, 0, sizeof A);
    A.in = x; A.gauge = ctx->gauge; A.g = g; A.kind = op->kind; A.dagger = dagger; A.hout = O;
    A.st = ctx->red.st; A.use_state = fuse ? fuse->use_state : 0;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_int, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_int, 0));
    halo_pack_kernel<<<ncta, 128, 0, ctx->stream2>>>(A);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_pack, ctx->stream2));
    if (op->kind == LQCD_WILSON) LQCD_TRY(launch_wilson_dslash(ctx, op, y, x, dagger, fuse, ctx->stream, &H));
    else                         LQCD_TRY(launch_staggered_dslash(ctx, op, y, x, dagger, fuse, ctx->stream, &H));
    // x must not be overwritten by later main-stream kernels before the pack has read it
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pack, 0));
    return LQCD_OK;
}

// LQCD_COMM_TIMING=1: per-application phase stamps (%globaltimer, ns) written by the kernels themselves -- pack CTAs, interior
// tiles, face tiles, time spent waiting for the neighbours' flags -- averaged over the recorded applications.  No host
// synchronisation is added, so the timeline is the one of the timed run.
int comm_timing_report(lqcd_ctx *ctx, const char *what) {
    CommState *c = ctx->comm;
    if (!c || !c->timing) return LQCD_OK;
    std::vector<unsigned long long> h(8 * LQCD_TIMING_SLOTS);
    CUDA_TRY(ctx, cudaMemcpy(h.data(), c->timing, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double a[7] = {0, 0, 0, 0, 0, 0, 0}, period = 0.0;
    int n = 0;
    unsigned long long prev_end = 0;
    const long long used = (long long)(c->halo_seq - c->timing_first + 1);
    const int last = (int)(used < LQCD_TIMING_SLOTS ? used : LQCD_TIMING_SLOTS);
    for (int i = last / 2; i < last; i++) {                      // second half: steady state
        const unsigned long long *s = &h[(size_t)i * 8];
        if (s[2] == ~0ull && s[4] == ~0ull) continue;
        unsigned long long t0 = s[0] < s[2] ? s[0] : s[2];
        if (s[4] < t0) t0 = s[4];
        const unsigned long long end = s[5] > s[3] ? s[5] : s[3];
        if (s[0] != ~0ull) { a[0] += (double)(s[0] - t0); a[1] += (double)(s[1] - t0); }
        if (s[2] != ~0ull) a[2] += (double)(s[3] - t0);
        if (s[4] != ~0ull) { a[3] += (double)(s[4] - t0); a[4] += (double)(s[5] - t0); }
        a[5] += (double)s[6]; a[6] += (double)s[7];
        if (prev_end && end > prev_end) period += (double)(end - prev_end);
        prev_end = end;
        n++;
    }
    if (n > 1)
        fprintf(stderr, "[lqcd comm timeline rank %d, %s, mean of %d applications, us after the first CTA started] pack %.1f..%.1f | interior tiles end %.1f | "
                        "face tiles %.1f..%.1f | flag wait: max %.1f, sum over face CTAs %.1f | period %.1f\n",
                ctx->rank, what, n, 1e-3 * a[0] / n, 1e-3 * a[1] / n, 1e-3 * a[2] / n, 1e-3 * a[3] / n, 1e-3 * a[4] / n, 1e-3 * a[6] / n, 1e-3 * a[5] / n,
                1e-3 * period / (n - 1));
    cudaFree(c->timing);
    c->timing = nullptr;
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
