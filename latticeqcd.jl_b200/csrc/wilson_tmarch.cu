// wilson_tmarch.cu -- Wilson Dslash, fp64, sm_100a: t-marching kernels with ALL compulsory traffic staged through shared memory by
// TMA bulk copies (cp.async.bulk + mbarrier).  EXPERIMENTAL, not the default: correct on hardware and they move the bytes they were
// designed to move, but both lose to the register-resident kernel (173 us per 32^4 application):
//   LQCD_WILSON_KERNEL=4  first generation, one CTA per SM, 144 KB of staging: 300 us (parity tests on 1 / 2 / 8 GPUs);
//   LQCD_WILSON_KERNEL=5  second generation (further down), two CTAs per SM, 96 KB of staging each: 182-184 us, bit-identical to
//                         the default kernel (single GPU on hardware; multi-rank under tests/emu only).
// Regular geometries only (x-line blocks, 4-warp (y,z) patches); launch_wilson_tmarch returns LQCD_ERR_STATE for anything else and
// the caller takes the register-resident kernel.
//
// Why (measured, profiles/r2a_*): the register-resident kernel moves 2.29 GB per 32^4 application from L2 to the SMs (2.2 KB/site,
// L1 hit rate 22 %) at 11 TB/s -- that is the L2->SM fabric limit (probe: 10.8 TB/s), so it sits at 0.76 of the HBM roofline and no
// occupancy variant (12 / 14 / 16 warps per SM) moves it.  A first t-marching kernel (round 1, spinor window only, links by LDG,
// 12 warps/SM) cut the fabric traffic but ran at 383 us: every step waited on link LDGs.  Here
//   * a CTA (4 warps = a 2x2 patch of 32-site blocks in (y,z)) marches along t, one persistent CTA per SM;
//   * the spinor records of the patch for slices t-1, t, t+1 live in a 3-slot shared-memory window: the own spinor, both x
//     neighbours, the in-patch y / z neighbours and both t neighbours are served from it (7 of 9 uses), each record is fetched
//     from L2 once per chunk step by ONE 6 KB bulk copy issued a full step ahead;
//   * the forward-link records of the patch (4 x 4.5 KB per block) live in four "planes", one per direction, refilled for the
//     next slice as soon as the whole CTA is done with the direction: the forward hop, the in-patch backward hops (the
//     neighbour's forward link) and the t-backward hop (the plane still holds slice t-1 when the step starts) read them from
//     shared memory, so every link enters the SM once;
//   * only the out-of-patch y / z neighbours (2 of 8 hops per site for a 2x2 patch) and their backward links use LDG.
// L2 -> SM traffic: 24 KB (compulsory) + ~16.5 KB per block-step = 1.27 KB/site instead of 2.2 KB; bytes in flight per SM are
// set by the copy schedule (~90 KB), not by registers or occupancy.
//
// Measured (B200, 32^4, profiles/r2b_wilson_tmarch_ncu_full.csv): 298-300 us; L2 -> SM traffic 1.43 GB (register kernel: 2.29 GB
// with full links, 2.03 GB with two-row links), DRAM 1.10 GB; but 144 KB of staging per CTA leaves ONE 4-warp CTA per SM = one warp
// per scheduler, and the issue slots are 20 % busy: stall samples 33 % fixed-latency (FP64 dependency chains), 22 % LDG of the
// out-of-patch neighbours, 14 % LDS latency, 7 % the four CTA barriers per step.  What it would take to win is in DESIGN.md section 8.
//
// Reference semantics: LinearAlgebra.mul!(y, D, x) / mul!(y, D', x) of LatticeDiracOperators.jl (upstream Wx!/Wdagx!, SURVEY.md
// App. C.1); call sites src/md/AbstractMD.jl:129, src/updates/standardHMC.jl:69-71, measure_Pion_correlator.jl:379,399.
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "wilson_spin.cuh"
#include "bulk_copy.cuh"
#include "site_map.cuh"
#include "link_load.cuh"
#ifdef TM_DEBUG
#include "../../tools/debug/tm_debug.cuh"
#endif
#include <cstdio>
#include <cstdlib>

#define TM_W 4                              // warps per CTA = blocks per patch
#define TM_REC (12 * 32)                    // complex numbers per spinor record (one 32-site block)
#define TM_REC_BYTES (TM_REC * 16)
#define TM_SUB (9 * 32)                     // complex numbers per link sub-record (one direction of one block)
#define TM_SUB_BYTES (TM_SUB * 16)
#define TM_SMEM_BYTES ((3 * TM_W * TM_REC + 4 * TM_W * TM_SUB) * 16 + 64)

struct TMArgs {
    WilsonArgs A;
    int Lc, nchunk, nsb, ntasks;            // t-steps per task, chunks per patch, blocks per t-slice, tasks = patches * chunks
    int prefetch;                           // second generation: pull the next task's first records into L2 during the last step
};

// hop arithmetic on operands already in registers: p = neighbour spinor, u = link (row-major 3x3)
template <int MU, int FWD, int DAG>
__device__ __forceinline__ void hop_regs(cplx (&acc)[12], const cplx (&p)[12], const cplx (&u)[9], bool wrapped, double phase) {
    constexpr int S = (FWD ^ DAG) ? -1 : +1;
    cplx h0[3], h1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) project<MU, S>(h0[c], h1[c], p[c], p[3 + c], p[6 + c], p[9 + c]);
    if (wrapped) {
#pragma unroll
        for (int c = 0; c < 3; c++) { h0[c] = cscale(phase, h0[c]); h1[c] = cscale(phase, h1[c]); }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        cplx g0 = cmake(0.0, 0.0), g1 = cmake(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            if (FWD) { cfma(g0, u[a * 3 + b], h0[b]); cfma(g1, u[a * 3 + b], h1[b]); }
            else     { cfmac(g0, u[b * 3 + a], h0[b]); cfmac(g1, u[b * 3 + a], h1[b]); }
        }
        reconstruct<MU, S>(acc, a, g0, g1);
    }
}

__device__ __forceinline__ void ld_spinor_s(cplx (&p)[12], const cplx *s) {         // shared memory (window)
#pragma unroll
    for (int k = 0; k < 12; k++) p[k] = s[k * 32];
}
__device__ __forceinline__ void ld_spinor_g(cplx (&p)[12], const cplx *__restrict__ gp) {   // global (out-of-patch neighbour)
#pragma unroll
    for (int k = 0; k < 12; k++) p[k] = __ldg(gp + k * 32);
}
__device__ __forceinline__ void ld_link_s(cplx (&u)[9], const cplx *s) {
#pragma unroll
    for (int e = 0; e < 9; e++) u[e] = s[e * 32];
}
__device__ __forceinline__ void ld_link_g(cplx (&u)[9], const cplx *__restrict__ gp) {
#pragma unroll
    for (int e = 0; e < 9; e++) u[e] = __ldg(gp + e * 32);
}

// off-rank hop (multi-GPU): the neighbour rank's pack kernel delivered the spin-projected half spinor of face site f into our halo
// slot (forward hop: P psi(n+mu), the local link u = U_mu(n) is applied here; backward hop: U^dag P psi(n-mu), complete).
// Same arithmetic as halo_hop() of wilson_kernel.cuh.
template <int MU, int FWD, int DAG>
__device__ __forceinline__ void halo_hop_regs(cplx (&acc)[12], const WilsonArgs &A, int f, const cplx (&u)[9]) {
    constexpr int S = (FWD ^ DAG) ? -1 : +1;
    const cplx *src = A.halo.recv[MU][FWD] + (size_t)(f >> 5) * (6 * 32) + (f & 31);
    cplx h0[3], h1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { h0[c] = __ldcg(src + c * 32); h1[c] = __ldcg(src + (3 + c) * 32); }
    const double phase = FWD ? (A.halo.plast[MU] ? A.bc[MU] : 1.0) : (A.halo.pfirst[MU] ? A.bc[MU] : 1.0);
    if (phase != 1.0) {
#pragma unroll
        for (int c = 0; c < 3; c++) { h0[c] = cscale(phase, h0[c]); h1[c] = cscale(phase, h1[c]); }
    }
    if (FWD) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cplx g0 = cmake(0.0, 0.0), g1 = cmake(0.0, 0.0);
#pragma unroll
            for (int b = 0; b < 3; b++) { cfma(g0, u[a * 3 + b], h0[b]); cfma(g1, u[a * 3 + b], h1[b]); }
            reconstruct<MU, S>(acc, a, g0, g1);
        }
    } else {
#pragma unroll
        for (int a = 0; a < 3; a++) reconstruct<MU, S>(acc, a, h0[a], h1[a]);
    }
}

// MULTI = 1: one rank of a process grid.  Hops that leave the local lattice in a partitioned direction read the halo slots the
// neighbours' pack kernels fill over NVLink (comm.cu); a CTA waits for the sequence flags right before its first such hop.  With
// T partitioned the march is cyclic and starts at t = 1, so the two slices that need the t halos (t = T-1, then t = 0) come
// LAST in every patch and the NVLink transfer is hidden behind the T-2 interior slices.
template <int DAG, int MULTI>
__global__ void __launch_bounds__(128, 1) wilson_tmarch_kernel(const TMArgs K) {
    const WilsonArgs &A = K.A;
    if (A.fuse.use_state && A.red.st->done) return;
    extern __shared__ __align__(128) unsigned char tm_smem[];
    cplx *const win = reinterpret_cast<cplx *>(tm_smem);                     // [3][W][12][32]
    cplx *const plane = win + 3 * TM_W * TM_REC;                             // [4 (mu)][W][9][32]
    uint64_t *const wbar = reinterpret_cast<uint64_t *>(plane + 4 * TM_W * TM_SUB);   // [3] window slots
    uint64_t *const pbar = wbar + 3;                                         // [4] link planes
    const Geom &g = A.g;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nsb = K.nsb, T = g.T, Lc = K.Lc;
    const int npatch = g.nt[0] * g.nt[1] * g.nt[2];

    if (threadIdx.x == 0) {
        for (int j = 0; j < 3; j++) mbar_init(&wbar[j], TM_W);
        for (int j = 0; j < 4; j++) mbar_init(&pbar[j], TM_W);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t wph = 0, pph = 0;            // parity to wait for next, one bit per barrier (every thread waits on every completion)
#ifdef TM_DEBUG
    __shared__ int tm_abort;
    if (threadIdx.x == 0) tm_abort = 0;
    __syncthreads();
    int dbg_r = 0, dbg_task = 0;
    auto wait_w = [&](int j) { tm_wait_dbg(wbar, j, (wph >> j) & 1u, 100 + j, dbg_r, dbg_task, &tm_abort); wph ^= 1u << j; };
    auto wait_p = [&](int j) { tm_wait_dbg(wbar, 3 + j, (pph >> j) & 1u, 200 + j, dbg_r, dbg_task, &tm_abort); pph ^= 1u << j; };
#else
    auto wait_w = [&](int j) { mbar_wait(&wbar[j], (wph >> j) & 1u); wph ^= 1u << j; };
    auto wait_p = [&](int j) { mbar_wait(&pbar[j], (pph >> j) & 1u); pph ^= 1u << j; };
#endif

    const int w0 = w % g.c[0], w1 = (w / g.c[0]) % g.c[1], w2 = w / (g.c[0] * g.c[1]);
    const int tshift = (MULTI && g.part[3]) ? 1 : 0;
    bool halo_ready = false;              // CTA-uniform: the neighbours' flags of this application have been seen
    double red[3] = {0.0, 0.0, 0.0};
    const double mk = -A.kappa;
    const double malpha = A.fuse.axpy_r ? -A.red.st->alpha : 0.0;
    cplx *const dst = A.fuse.axpy_r ? A.fuse.axpy_r : A.out;

    for (int task = blockIdx.x; task < K.ntasks; task += gridDim.x) {
        // tasks are numbered chunk-major: the CTAs in flight sweep the lattice in t together, so the out-of-patch neighbours and
        // the next slice are L2 hits (each record comes from HBM once per application)
        const int patch = task % npatch, chunk = task / npatch;
        const int p0 = patch % g.nt[0], p1 = (patch / g.nt[0]) % g.nt[1], p2 = patch / (g.nt[0] * g.nt[1]);
        const int bslice = (p0 * g.c[0] + w0) + g.nb[0] * ((p1 * g.c[1] + w1) + g.nb[1] * (p2 * g.c[2] + w2));
        const int t0 = chunk * Lc + tshift;                                      // slice of step r: (t0 + r - 1) mod T
        auto slice_of = [&](int rel) { return (t0 - 1 + rel + T) % T; };         // rel = step index (0 = the slice before the first step)

        // t-invariant spatial neighbour tables of this lane: block in the slice, lane, patch position (-1 = out of patch), wrap
        const int ssl = bslice * 32 + lane;
        const int cx = ssl % g.X, cy = (ssl / g.X) % g.Y, cz = ssl / (g.X * g.Y);
        int nbl[6], nl[6], nw[6];
        bool wr[6];
        {
            const int coord[3] = {cx, cy, cz}, dim[3] = {g.X, g.Y, g.Z}, stride[3] = {1, g.X, g.X * g.Y};
#pragma unroll
            for (int mu = 0; mu < 3; mu++) {
#pragma unroll
                for (int f = 0; f < 2; f++) {                        // f = 0: forward (+mu), f = 1: backward (-mu)
                    const int d = mu * 2 + f;
                    const bool wrapd = f == 0 ? (coord[mu] == dim[mu] - 1) : (coord[mu] == 0);
                    const int nssl = f == 0 ? (wrapd ? ssl - (dim[mu] - 1) * stride[mu] : ssl + stride[mu])
                                            : (wrapd ? ssl + (dim[mu] - 1) * stride[mu] : ssl - stride[mu]);
                    wr[d] = wrapd;
                    nbl[d] = nssl >> 5; nl[d] = nssl & 31;
                    const int q0 = nbl[d] % g.nb[0], q1 = (nbl[d] / g.nb[0]) % g.nb[1], q2 = nbl[d] / (g.nb[0] * g.nb[1]);
                    const bool inp = q0 >= p0 * g.c[0] && q0 < (p0 + 1) * g.c[0] && q1 >= p1 * g.c[1] && q1 < (p1 + 1) * g.c[1] &&
                                     q2 >= p2 * g.c[2] && q2 < (p2 + 1) * g.c[2];
                    nw[d] = inp ? (q0 - p0 * g.c[0]) + g.c[0] * ((q1 - p1 * g.c[1]) + g.c[1] * (q2 - p2 * g.c[2])) : -1;
                }
            }
        }

        __syncthreads();                  // everybody is done reading the previous task's window and planes
        if (MULTI && !halo_ready) {       // a patch on a partitioned spatial face needs its halos from the first step on
            bool need = false;
            const int pc[3] = {p0, p1, p2};
            for (int mu = 0; mu < 3; mu++) need = need || (g.part[mu] && (pc[mu] == 0 || pc[mu] == g.nt[mu] - 1));
            if (need) { wait_halo_flags(g, A.halo); halo_ready = true; }
        }
        if (lane == 0) {
            for (int rel = 0; rel < 3; rel++) {                                  // slices t0-1, t0, t0+1 -> slots 0, 1, 2
                const int tt = slice_of(rel);
                mbar_arrive_expect_tx(&wbar[rel], TM_REC_BYTES);
                bulk_g2s(win + ((size_t)rel * TM_W + w) * TM_REC, A.in + ((size_t)bslice + (size_t)tt * nsb) * TM_REC, TM_REC_BYTES, &wbar[rel]);
            }
            const size_t b0 = (size_t)bslice + (size_t)slice_of(1) * nsb, bm = (size_t)bslice + (size_t)slice_of(0) * nsb;
            for (int mu = 0; mu < 3; mu++) {                                     // spatial forward links of slice t0
                mbar_arrive_expect_tx(&pbar[mu], TM_SUB_BYTES);
                bulk_g2s(plane + ((size_t)mu * TM_W + w) * TM_SUB, A.gauge + (b0 * 4 + mu) * TM_SUB, TM_SUB_BYTES, &pbar[mu]);
            }
            mbar_arrive_expect_tx(&pbar[3], TM_SUB_BYTES);                       // t links of slice t0-1 (backward hop of the first step)
            bulk_g2s(plane + ((size_t)3 * TM_W + w) * TM_SUB, A.gauge + (bm * 4 + 3) * TM_SUB, TM_SUB_BYTES, &pbar[3]);
        }

        for (int r = 1; r <= Lc; r++) {
            const int t = slice_of(r);
#ifdef TM_DEBUG
            dbg_r = r; dbg_task = task;
#endif
            const size_t blk = (size_t)bslice + (size_t)t * nsb;
            if (MULTI && !halo_ready && g.part[3] && (t == T - 1 || t == 0)) { wait_halo_flags(g, A.halo); halo_ready = true; }
            const cplx *cur = win + (size_t)(r % 3) * TM_W * TM_REC;
            const cplx *up = win + (size_t)((r + 1) % 3) * TM_W * TM_REC;
            const cplx *dn = win + (size_t)((r - 1) % 3) * TM_W * TM_REC;
            const size_t base = blk * TM_REC + lane;
            if (A.fuse.axpy_r || A.fuse.dot_with || A.fuse.shift_src) {          // epilogue operands: start them towards L2 now
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    if (A.fuse.axpy_r) prefetch_l2(A.fuse.axpy_r + base + k * 32);
                    if (A.fuse.dot_with) prefetch_l2(A.fuse.dot_with + base + k * 32);
                    if (A.fuse.shift_src) prefetch_l2(A.fuse.shift_src + base + k * 32);
                }
            }
            cplx acc[12];
#pragma unroll
            for (int k = 0; k < 12; k++) acc[k] = cmake(0.0, 0.0);
            cplx p[12], u[9];

            // ---- t backward: slice t-1 from the window, its t links still in plane 3 -------------------------------------------
            if (r == 1) { wait_w(0); wait_w(1); wait_p(3); }      // later steps: plane 3 was waited for by the previous step's forward hop
            // (accumulated apart and added LAST: the hop sum then has the order x+ x- y+ y- z+ z- t+ t- of the register-resident
            //  kernel, and the two kernel families, the multi-RHS and the slab-pipelined paths stay bit-identical)
            cplx acct[12];
#pragma unroll
            for (int k = 0; k < 12; k++) acct[k] = cmake(0.0, 0.0);
            if (MULTI && g.part[3] && t == 0) halo_hop_regs<3, 0, DAG>(acct, A, face_index<3>(g, cx, cy, cz, t), u);
            else {
                ld_spinor_s(p, dn + (size_t)w * TM_REC + lane);
                ld_link_s(u, plane + ((size_t)3 * TM_W + w) * TM_SUB + lane);
                hop_regs<3, 0, DAG>(acct, p, u, t == 0, A.bc[3]);
            }
            __syncthreads();
            if (lane == 0) {
                mbar_arrive_expect_tx(&pbar[3], TM_SUB_BYTES);                   // t links of slice t (forward hop at the end of the step)
                bulk_g2s(plane + ((size_t)3 * TM_W + w) * TM_SUB, A.gauge + (blk * 4 + 3) * TM_SUB, TM_SUB_BYTES, &pbar[3]);
                if (r + 2 <= Lc + 1) {                                           // slice t+2 into the slot slice t-1 just left
                    const int rel = r + 2, tt = slice_of(rel);
                    mbar_arrive_expect_tx(&wbar[rel % 3], TM_REC_BYTES);
                    bulk_g2s(win + ((size_t)(rel % 3) * TM_W + w) * TM_REC, A.in + ((size_t)bslice + (size_t)tt * nsb) * TM_REC, TM_REC_BYTES, &wbar[rel % 3]);
                }
            }

            // ---- spatial directions: plane mu holds the forward links of slice t for the whole patch ---------------------------
#define TM_SPATIAL(MU)                                                                                                      \
            {                                                                                                               \
                wait_p(MU);                                                                                                 \
                const int df = MU * 2, db = MU * 2 + 1;                                                                     \
                const cplx *pl = plane + (size_t)MU * TM_W * TM_SUB;                                                        \
                ld_link_s(u, pl + (size_t)w * TM_SUB + lane);                                                               \
                if (MULTI && wr[df] && g.part[MU]) halo_hop_regs<MU, 1, DAG>(acc, A, face_index<MU>(g, cx, cy, cz, t), u);  \
                else {                                                                                                      \
                    if (nw[df] >= 0) ld_spinor_s(p, cur + (size_t)nw[df] * TM_REC + nl[df]);                                \
                    else             ld_spinor_g(p, A.in + ((size_t)nbl[df] + (size_t)t * nsb) * TM_REC + nl[df]);          \
                    hop_regs<MU, 1, DAG>(acc, p, u, wr[df], A.bc[MU]);                                                      \
                }                                                                                                           \
                if (MULTI && wr[db] && g.part[MU]) halo_hop_regs<MU, 0, DAG>(acc, A, face_index<MU>(g, cx, cy, cz, t), u);  \
                else if (nw[db] >= 0) { ld_spinor_s(p, cur + (size_t)nw[db] * TM_REC + nl[db]); ld_link_s(u, pl + (size_t)nw[db] * TM_SUB + nl[db]); } \
                else {                                                                                                      \
                    ld_spinor_g(p, A.in + ((size_t)nbl[db] + (size_t)t * nsb) * TM_REC + nl[db]);                           \
                    ld_link_g(u, A.gauge + (((size_t)nbl[db] + (size_t)t * nsb) * 4 + MU) * TM_SUB + nl[db]);               \
                }                                                                                                           \
                if (!(MULTI && wr[db] && g.part[MU])) hop_regs<MU, 0, DAG>(acc, p, u, wr[db], A.bc[MU]);                    \
                __syncthreads();                                                                                            \
                if (lane == 0 && r < Lc) {                                       /* forward links of slice t+1 */           \
                    mbar_arrive_expect_tx(&pbar[MU], TM_SUB_BYTES);                                                         \
                    bulk_g2s(plane + ((size_t)MU * TM_W + w) * TM_SUB, A.gauge + (((size_t)bslice + (size_t)slice_of(r + 1) * nsb) * 4 + MU) * TM_SUB, TM_SUB_BYTES, &pbar[MU]); \
                }                                                                                                           \
            }
            TM_SPATIAL(0) TM_SPATIAL(1) TM_SPATIAL(2)
#undef TM_SPATIAL

            // ---- t forward: slice t+1 from the window, t links of slice t (requested after the backward hop) -------------------
            wait_w((r + 1) % 3);
            wait_p(3);
            ld_link_s(u, plane + ((size_t)3 * TM_W + w) * TM_SUB + lane);
            if (MULTI && g.part[3] && t == T - 1) halo_hop_regs<3, 1, DAG>(acc, A, face_index<3>(g, cx, cy, cz, t), u);
            else {
                ld_spinor_s(p, up + (size_t)w * TM_REC + lane);
                hop_regs<3, 1, DAG>(acc, p, u, t == T - 1, A.bc[3]);
            }
#pragma unroll
            for (int k = 0; k < 12; k++) acc[k] = cadd(acc[k], acct[k]);

            // ---- epilogue: y = x - kappa * hops (+ fused shift / CG residual update / reductions), as wilson_kernel.cuh ----------
            const cplx *own = cur + (size_t)w * TM_REC + lane;
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const cplx xi = own[k * 32];
                cplx yk = cmake(fma(mk, acc[k].x, xi.x), fma(mk, acc[k].y, xi.y));
                if (A.fuse.shift_src) {
                    const cplx sv = ldg128(A.fuse.shift_src + base + k * 32);
                    yk.x = fma(A.fuse.shift, sv.x, yk.x); yk.y = fma(A.fuse.shift, sv.y, yk.y);
                }
                if (A.fuse.axpy_r) {
                    const cplx rv = A.fuse.axpy_r[base + k * 32];
                    yk = cmake(fma(malpha, yk.x, rv.x), fma(malpha, yk.y, rv.y));
                }
                if (A.fuse.dot_with) {
                    const cplx wv = ldg128(A.fuse.dot_with + base + k * 32);
                    red[0] = fma(wv.x, yk.x, red[0]); red[0] = fma(wv.y, yk.y, red[0]);
                    red[1] = fma(wv.x, yk.y, red[1]); red[1] = fma(-wv.y, yk.x, red[1]);
                }
                red[2] = fma(yk.x, yk.x, red[2]); red[2] = fma(yk.y, yk.y, red[2]);
                dst[base + k * 32] = yk;
            }
        }
    }
    if (A.fuse.dot_with || A.fuse.want_norm) grid_reduce_finish<3>(red, A.red, A.fuse.finish);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Second generation (LQCD_WILSON_KERNEL=5): the same march with HALF the staging, so that TWO CTAs (8 warps, two per scheduler) share
// an SM -- the first kernel's one warp per scheduler is what left its issue slots 80 % idle.
//   * link planes hold two-row records (links12.cu: 3 KB instead of 4.5 KB per block and direction; the third row is rebuilt in
//     registers by the same link_row3() as the register-resident kernel, so the two stay bit-identical);
//   * the window has TWO slots (t, t+1): the t-backward hop of slice t+1, U_t(t)^dag P psi(t), needs only operands the step of
//     slice t already holds (own spinor, own t link), so it is computed there and CARRIED to the next step in 24 registers; the
//     first step of a task computes its carry from 18 global loads;
//   * three CTA barriers per step (one per spatial direction, after which the direction's plane is refilled); the window slot and
//     the t plane are private to a warp once the z phase is over and are refilled after a warp-level sync.
// 96 KB of staging per CTA.  Needs SU(3) links (two-row copy valid); otherwise the caller takes the register-resident kernel.
//
// Measured (B200, 32^4, profiles/r2j_wilson_tmarch2.txt): 184.1 us with the automatic chunking (8 chunks of 4 slices: 2048 tasks on
// 296 resident CTAs = 6.9 rounds), 191 / 196 / 199 / 197 us with 1 / 2 / 4 / 16 chunks.  A fit of those five numbers gives 5.8 us
// per step and ~0.5 step of pipeline fill per task, i.e. ~161 us if the 256 patches could be spread evenly over 296 CTA slots with
// long chunks -- they cannot (256 x nchunk tasks), and splitting the (patch, t) sequence evenly would give up the common t wavefront
// that keeps the out-of-patch neighbours L2 hits.  Pulling the next task's first records into L2 during the last step
// (LQCD_TM_PREFETCH=1) did not help: 187.9 us.  So: 1.63x faster than the first generation, 6 % slower than the default kernel.
#define TM2_SUB (6 * 32)
#define TM2_SUB_BYTES (TM2_SUB * 16)
#define TM2_SMEM_BYTES ((2 * TM_W * TM_REC + 4 * TM_W * TM2_SUB) * 16 + 64)

__device__ __forceinline__ void ld_link12_s(cplx (&u)[9], const cplx *s) {
#pragma unroll
    for (int e = 0; e < 6; e++) u[e] = s[e * 32];
    link_row3(u);
}
__device__ __forceinline__ void ld_link12_g(cplx (&u)[9], const cplx *__restrict__ gp) {
#pragma unroll
    for (int e = 0; e < 6; e++) u[e] = __ldg(gp + e * 32);
    link_row3(u);
}
// the t-backward hop of the NEXT slice from this site's spinor and t link: g = U_t(n)^dag P psi(n) (times the boundary phase when
// the next slice is t = 0); reconstruct<3, S>(acc, a, g0[a], g1[a]) there completes hop_regs<3, 0, DAG> bit for bit
template <int DAG>
__device__ __forceinline__ void carry_make(cplx (&c0)[3], cplx (&c1)[3], const cplx (&p)[12], const cplx (&u)[9], bool wrapped, double phase) {
    constexpr int S = DAG ? -1 : +1;
    cplx h0[3], h1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) project<3, S>(h0[c], h1[c], p[c], p[3 + c], p[6 + c], p[9 + c]);
    if (wrapped) {
#pragma unroll
        for (int c = 0; c < 3; c++) { h0[c] = cscale(phase, h0[c]); h1[c] = cscale(phase, h1[c]); }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        cplx g0 = cmake(0.0, 0.0), g1 = cmake(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < 3; b++) { cfmac(g0, u[b * 3 + a], h0[b]); cfmac(g1, u[b * 3 + a], h1[b]); }
        c0[a] = g0; c1[a] = g1;
    }
}

// PIPE = 1 (LQCD_TM_PIPE=1, not yet on hardware): tasks are pipelined -- during the LAST step of a task every plane is refilled with
// the NEXT task's first slice as soon as its direction is done, and every warp requests its own window records and t plane of the
// next task right after its epilogue, so a task starts with its copies a step old instead of just issued (no CTA barrier between
// tasks: the mbarriers order the cross-warp reads); and when a block spans the whole x extent the x plane is private to a warp,
// so the x phase ends with a warp-level sync (two CTA barriers per step).
template <int DAG, int MULTI, int PIPE>
__global__ void __launch_bounds__(128, 2) wilson_tmarch2_kernel(const TMArgs K) {
    const WilsonArgs &A = K.A;
    if (A.fuse.use_state && A.red.st->done) return;
    extern __shared__ __align__(128) unsigned char tm_smem[];
    cplx *const win = reinterpret_cast<cplx *>(tm_smem);                     // [2][W][12][32]
    cplx *const plane = win + 2 * TM_W * TM_REC;                             // [4 (mu)][W][6][32]
    uint64_t *const wbar = reinterpret_cast<uint64_t *>(plane + 4 * TM_W * TM2_SUB);  // [2] window slots
    uint64_t *const pbar = wbar + 2;                                         // [4] link planes
    const Geom &g = A.g;
    const cplx *const links = A.links12;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nsb = K.nsb, T = g.T, Lc = K.Lc;
    const int npatch = g.nt[0] * g.nt[1] * g.nt[2];

    if (threadIdx.x == 0) {
        for (int j = 0; j < 2; j++) mbar_init(&wbar[j], TM_W);
        for (int j = 0; j < 4; j++) mbar_init(&pbar[j], TM_W);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t wph = 0, pph = 0;
    auto wait_w = [&](int j) { mbar_wait(&wbar[j], (wph >> j) & 1u); wph ^= 1u << j; };
    auto wait_p = [&](int j) { mbar_wait(&pbar[j], (pph >> j) & 1u); pph ^= 1u << j; };

    const int w0 = w % g.c[0], w1 = (w / g.c[0]) % g.c[1], w2 = w / (g.c[0] * g.c[1]);
    const int tshift = (MULTI && g.part[3]) ? 1 : 0;
    bool halo_ready = false;
    double red[3] = {0.0, 0.0, 0.0};
    const double mk = -A.kappa;
    const double malpha = A.fuse.axpy_r ? -A.red.st->alpha : 0.0;
    cplx *const dst = A.fuse.axpy_r ? A.fuse.axpy_r : A.out;
    const bool xpriv = PIPE && g.nb[0] == 1;          // both x neighbours of a site live in the site's own block
    bool primed = false;                              // PIPE: the previous task's last step requested this task's first copies

    for (int task = blockIdx.x; task < K.ntasks; task += gridDim.x) {
        const int patch = task % npatch, chunk = task / npatch;
        const int p0 = patch % g.nt[0], p1 = (patch / g.nt[0]) % g.nt[1], p2 = patch / (g.nt[0] * g.nt[1]);
        const int bslice = (p0 * g.c[0] + w0) + g.nb[0] * ((p1 * g.c[1] + w1) + g.nb[1] * (p2 * g.c[2] + w2));
        const int t0 = chunk * Lc + tshift;
        auto slice_of = [&](int rel) { return (t0 - 1 + rel + T) % T; };

        const int ssl = bslice * 32 + lane;
        const int cx = ssl % g.X, cy = (ssl / g.X) % g.Y, cz = ssl / (g.X * g.Y);
        int nbl[6], nl[6], nw[6];
        bool wr[6];
        {
            const int coord[3] = {cx, cy, cz}, dim[3] = {g.X, g.Y, g.Z}, stride[3] = {1, g.X, g.X * g.Y};
#pragma unroll
            for (int mu = 0; mu < 3; mu++) {
#pragma unroll
                for (int f = 0; f < 2; f++) {
                    const int d = mu * 2 + f;
                    const bool wrapd = f == 0 ? (coord[mu] == dim[mu] - 1) : (coord[mu] == 0);
                    const int nssl = f == 0 ? (wrapd ? ssl - (dim[mu] - 1) * stride[mu] : ssl + stride[mu])
                                            : (wrapd ? ssl + (dim[mu] - 1) * stride[mu] : ssl - stride[mu]);
                    wr[d] = wrapd;
                    nbl[d] = nssl >> 5; nl[d] = nssl & 31;
                    const int q0 = nbl[d] % g.nb[0], q1 = (nbl[d] / g.nb[0]) % g.nb[1], q2 = nbl[d] / (g.nb[0] * g.nb[1]);
                    const bool inp = q0 >= p0 * g.c[0] && q0 < (p0 + 1) * g.c[0] && q1 >= p1 * g.c[1] && q1 < (p1 + 1) * g.c[1] &&
                                     q2 >= p2 * g.c[2] && q2 < (p2 + 1) * g.c[2];
                    nw[d] = inp ? (q0 - p0 * g.c[0]) + g.c[0] * ((q1 - p1 * g.c[1]) + g.c[1] * (q2 - p2 * g.c[2])) : -1;
                }
            }
        }

        if (!(PIPE && primed)) __syncthreads();      // everybody is done reading the previous task's window and planes
        if (MULTI && !halo_ready) {
            bool need = false;
            const int pc[3] = {p0, p1, p2};
            for (int mu = 0; mu < 3; mu++) need = need || (g.part[mu] && (pc[mu] == 0 || pc[mu] == g.nt[mu] - 1));
            if (need) { wait_halo_flags(g, A.halo); halo_ready = true; }
        }
        if (lane == 0 && !(PIPE && primed)) {
            for (int rel = 1; rel <= 2; rel++) {                                 // slices of step 1 and step 2 -> slots 1, 0
                mbar_arrive_expect_tx(&wbar[rel & 1], TM_REC_BYTES);
                bulk_g2s(win + ((size_t)(rel & 1) * TM_W + w) * TM_REC, A.in + ((size_t)bslice + (size_t)slice_of(rel) * nsb) * TM_REC, TM_REC_BYTES, &wbar[rel & 1]);
            }
            const size_t b0 = (size_t)bslice + (size_t)slice_of(1) * nsb;
            for (int mu = 0; mu < 4; mu++) {                                     // forward links of the first slice
                mbar_arrive_expect_tx(&pbar[mu], TM2_SUB_BYTES);
                bulk_g2s(plane + ((size_t)mu * TM_W + w) * TM2_SUB, links + (b0 * 4 + mu) * TM2_SUB, TM2_SUB_BYTES, &pbar[mu]);
            }
        }
        cplx c0[3], c1[3];                // carried t-backward hop
        {
            const int tf = slice_of(1);
            if (!(MULTI && g.part[3] && tf == 0)) {
                const size_t bm = (size_t)bslice + (size_t)slice_of(0) * nsb;
                cplx p[12], u[9];
                ld_spinor_g(p, A.in + bm * TM_REC + lane);
                ld_link12_g(u, links + (bm * 4 + 3) * TM2_SUB + lane);
                carry_make<DAG>(c0, c1, p, u, tf == 0, A.bc[3]);
            } else {
#pragma unroll
                for (int a = 0; a < 3; a++) { c0[a] = cmake(0.0, 0.0); c1[a] = cmake(0.0, 0.0); }
            }
        }

        // PIPE: this warp's block and first slices of the CTA's next task
        const bool nxt = PIPE && task + (int)gridDim.x < K.ntasks;
        size_t nbs = 0;
        int nt0 = 0;
        if (nxt) {
            const int ntask = task + (int)gridDim.x, np = ntask % npatch;
            const int n0 = np % g.nt[0], n1 = (np / g.nt[0]) % g.nt[1], n2 = np / (g.nt[0] * g.nt[1]);
            nbs = (size_t)((n0 * g.c[0] + w0) + g.nb[0] * ((n1 * g.c[1] + w1) + g.nb[1] * (n2 * g.c[2] + w2)));
            nt0 = (ntask / npatch) * Lc + tshift;
        }
        auto nslice = [&](int rel) { return (nt0 - 1 + rel + T) % T; };

        for (int r = 1; r <= Lc; r++) {
            const int t = slice_of(r);
            const size_t blk = (size_t)bslice + (size_t)t * nsb;
            if (MULTI && !halo_ready && g.part[3] && (t == T - 1 || t == 0)) { wait_halo_flags(g, A.halo); halo_ready = true; }
            const cplx *cur = win + (size_t)(r & 1) * TM_W * TM_REC;
            const cplx *up = win + (size_t)((r + 1) & 1) * TM_W * TM_REC;
            const size_t base = blk * TM_REC + lane;
            if (A.fuse.axpy_r || A.fuse.dot_with || A.fuse.shift_src) {
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    if (A.fuse.axpy_r) prefetch_l2(A.fuse.axpy_r + base + k * 32);
                    if (A.fuse.dot_with) prefetch_l2(A.fuse.dot_with + base + k * 32);
                    if (A.fuse.shift_src) prefetch_l2(A.fuse.shift_src + base + k * 32);
                }
            }
            if (K.prefetch && r == Lc && task + (int)gridDim.x < K.ntasks) {
                // the next task starts cold (a new chunk of another patch): its window records, link planes and carry operands are
                // on their way to L2 while this task's last step runs, so its priming copies and loads are L2 hits
                const int ntask = task + (int)gridDim.x, np = ntask % npatch;
                const int n0 = np % g.nt[0], n1 = (np / g.nt[0]) % g.nt[1], n2 = np / (g.nt[0] * g.nt[1]);
                const size_t nbs = (size_t)((n0 * g.c[0] + w0) + g.nb[0] * ((n1 * g.c[1] + w1) + g.nb[1] * (n2 * g.c[2] + w2)));
                const int nt0 = (ntask / npatch) * Lc + tshift;
                for (int rel = 0; rel < 3; rel++) {
                    const size_t nb = nbs + (size_t)((nt0 - 1 + rel + T) % T) * nsb;
                    const char *rec = reinterpret_cast<const char *>(A.in + nb * TM_REC);
                    prefetch_l2(rec + lane * 128);
                    if (lane < 16) prefetch_l2(rec + (32 + lane) * 128);
                    if (lane < 24) {
                        if (rel == 0) prefetch_l2(reinterpret_cast<const char *>(links + (nb * 4 + 3) * TM2_SUB) + lane * 128);
                        if (rel == 1) {
                            for (int mu = 0; mu < 4; mu++) prefetch_l2(reinterpret_cast<const char *>(links + (nb * 4 + mu) * TM2_SUB) + lane * 128);
                        }
                    }
                }
            }
            cplx acc[12];
#pragma unroll
            for (int k = 0; k < 12; k++) acc[k] = cmake(0.0, 0.0);
            cplx p[12], u[9];
            if (r == 1) wait_w(1);        // later steps: the slot was waited for as `up` by the previous step

#define TM2_SPATIAL(MU)                                                                                                     \
            {                                                                                                               \
                wait_p(MU);                                                                                                 \
                const int df = MU * 2, db = MU * 2 + 1;                                                                     \
                const cplx *pl = plane + (size_t)MU * TM_W * TM2_SUB;                                                       \
                ld_link12_s(u, pl + (size_t)w * TM2_SUB + lane);                                                            \
                if (MULTI && wr[df] && g.part[MU]) halo_hop_regs<MU, 1, DAG>(acc, A, face_index<MU>(g, cx, cy, cz, t), u);  \
                else {                                                                                                      \
                    if (nw[df] >= 0) ld_spinor_s(p, cur + (size_t)nw[df] * TM_REC + nl[df]);                                \
                    else             ld_spinor_g(p, A.in + ((size_t)nbl[df] + (size_t)t * nsb) * TM_REC + nl[df]);          \
                    hop_regs<MU, 1, DAG>(acc, p, u, wr[df], A.bc[MU]);                                                      \
                }                                                                                                           \
                if (MULTI && wr[db] && g.part[MU]) halo_hop_regs<MU, 0, DAG>(acc, A, face_index<MU>(g, cx, cy, cz, t), u);  \
                else if (nw[db] >= 0) { ld_spinor_s(p, cur + (size_t)nw[db] * TM_REC + nl[db]); ld_link12_s(u, pl + (size_t)nw[db] * TM2_SUB + nl[db]); } \
                else {                                                                                                      \
                    ld_spinor_g(p, A.in + ((size_t)nbl[db] + (size_t)t * nsb) * TM_REC + nl[db]);                           \
                    ld_link12_g(u, links + (((size_t)nbl[db] + (size_t)t * nsb) * 4 + MU) * TM2_SUB + nl[db]);              \
                }                                                                                                           \
                if (!(MULTI && wr[db] && g.part[MU])) hop_regs<MU, 0, DAG>(acc, p, u, wr[db], A.bc[MU]);                    \
                if (MU == 0 && xpriv) __syncwarp(); else __syncthreads();                                                   \
                if (lane == 0 && (r < Lc || nxt)) {                              /* forward links of the next slice (PIPE, last step: of the next task's first slice) */ \
                    const size_t nb_ = r < Lc ? (size_t)bslice + (size_t)slice_of(r + 1) * nsb : nbs + (size_t)nslice(1) * nsb;   \
                    mbar_arrive_expect_tx(&pbar[MU], TM2_SUB_BYTES);                                                        \
                    bulk_g2s(plane + ((size_t)MU * TM_W + w) * TM2_SUB, links + (nb_ * 4 + MU) * TM2_SUB, TM2_SUB_BYTES, &pbar[MU]); \
                }                                                                                                           \
            }
            TM2_SPATIAL(0) TM2_SPATIAL(1) TM2_SPATIAL(2)
#undef TM2_SPATIAL
            // from here on the warp touches only its own records (cur[w], up[w], plane 3 [w])

            // ---- t forward ---------------------------------------------------------------------------------------------------------
            wait_w((r + 1) & 1);
            wait_p(3);
            ld_link12_s(u, plane + ((size_t)3 * TM_W + w) * TM2_SUB + lane);
            if (MULTI && g.part[3] && t == T - 1) halo_hop_regs<3, 1, DAG>(acc, A, face_index<3>(g, cx, cy, cz, t), u);
            else {
                ld_spinor_s(p, up + (size_t)w * TM_REC + lane);
                hop_regs<3, 1, DAG>(acc, p, u, t == T - 1, A.bc[3]);
            }
            // ---- t backward: the carry of the previous step (or the halo) ---------------------------------------------------------
            if (MULTI && g.part[3] && t == 0) halo_hop_regs<3, 0, DAG>(acc, A, face_index<3>(g, cx, cy, cz, t), u);
            else {
#pragma unroll
                for (int a = 0; a < 3; a++) reconstruct<3, (DAG ? -1 : +1)>(acc, a, c0[a], c1[a]);
            }
            ld_spinor_s(p, cur + (size_t)w * TM_REC + lane);                     // own spinor: next carry and the epilogue
            if (r < Lc && !(MULTI && g.part[3] && t == T - 1)) carry_make<DAG>(c0, c1, p, u, t == T - 1, A.bc[3]);

            // ---- epilogue -----------------------------------------------------------------------------------------------------------
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const cplx xi = p[k];
                cplx yk = cmake(fma(mk, acc[k].x, xi.x), fma(mk, acc[k].y, xi.y));
                if (A.fuse.shift_src) {
                    const cplx sv = ldg128(A.fuse.shift_src + base + k * 32);
                    yk.x = fma(A.fuse.shift, sv.x, yk.x); yk.y = fma(A.fuse.shift, sv.y, yk.y);
                }
                if (A.fuse.axpy_r) {
                    const cplx rv = A.fuse.axpy_r[base + k * 32];
                    yk = cmake(fma(malpha, yk.x, rv.x), fma(malpha, yk.y, rv.y));
                }
                if (A.fuse.dot_with) {
                    const cplx wv = ldg128(A.fuse.dot_with + base + k * 32);
                    red[0] = fma(wv.x, yk.x, red[0]); red[0] = fma(wv.y, yk.y, red[0]);
                    red[1] = fma(wv.x, yk.y, red[1]); red[1] = fma(-wv.y, yk.x, red[1]);
                }
                red[2] = fma(yk.x, yk.x, red[2]); red[2] = fma(yk.y, yk.y, red[2]);
                dst[base + k * 32] = yk;
            }
            __syncwarp();                 // every lane has consumed its own spinor and t link: the warp's records may be overwritten
            if (lane == 0) {
                if (r < Lc) {                                                    // t links of the next slice
                    mbar_arrive_expect_tx(&pbar[3], TM2_SUB_BYTES);
                    bulk_g2s(plane + ((size_t)3 * TM_W + w) * TM2_SUB, links + (((size_t)bslice + (size_t)slice_of(r + 1) * nsb) * 4 + 3) * TM2_SUB, TM2_SUB_BYTES, &pbar[3]);
                }
                if (r + 2 <= Lc + 1) {                                           // slice of step r+2 into the slot this step leaves
                    mbar_arrive_expect_tx(&wbar[r & 1], TM_REC_BYTES);
                    bulk_g2s(win + ((size_t)(r & 1) * TM_W + w) * TM_REC, A.in + ((size_t)bslice + (size_t)slice_of(r + 2) * nsb) * TM_REC, TM_REC_BYTES, &wbar[r & 1]);
                }
                if (r == Lc && nxt) {                                            // PIPE: the next task's t plane and both window records of this warp
                    const size_t nb1 = nbs + (size_t)nslice(1) * nsb;
                    mbar_arrive_expect_tx(&pbar[3], TM2_SUB_BYTES);
                    bulk_g2s(plane + ((size_t)3 * TM_W + w) * TM2_SUB, links + (nb1 * 4 + 3) * TM2_SUB, TM2_SUB_BYTES, &pbar[3]);
                    for (int rel = 1; rel <= 2; rel++) {
                        mbar_arrive_expect_tx(&wbar[rel & 1], TM_REC_BYTES);
                        bulk_g2s(win + ((size_t)(rel & 1) * TM_W + w) * TM_REC, A.in + (nbs + (size_t)nslice(rel) * nsb) * TM_REC, TM_REC_BYTES, &wbar[rel & 1]);
                    }
                }
            }
        }
        primed = nxt;
    }
    if (A.fuse.dot_with || A.fuse.want_norm) grid_reduce_finish<3>(red, A.red, A.fuse.finish);
}

// geometry / configuration test shared with comm.cu (which then feeds the halo slots with the separate pack kernel)
bool wilson_tmarch_ok(const lqcd_ctx *ctx, const lqcd_op *op) {
    static int family = -1;
    if (family < 0) { const char *e = getenv("LQCD_WILSON_KERNEL"); family = e ? atoi(e) : 0; }
    const Geom &g = ctx->g;
    if ((family != 4 && family != 5) || op->kind != LQCD_WILSON || op->csw != 0.0 || op->r != 1.0) return false;
    return g.regular && g.s[3] == 1 && g.c[3] == 1 && g.c[0] * g.c[1] * g.c[2] == TM_W && g.T >= 2 && (!g.part[3] || g.T >= 3);
}

// LQCD_OK if launched; LQCD_ERR_STATE if the geometry / variant does not qualify (caller falls back to the register-resident kernel)
int ensure_links12(lqcd_ctx *ctx, int *use);       // links12.cu

int launch_wilson_tmarch(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, cudaStream_t s, bool halo, bool self_pack) {
    const Geom &g = ctx->g;
    static int family = -1;
    if (family < 0) { const char *e = getenv("LQCD_WILSON_KERNEL"); family = e ? atoi(e) : 0; }
    const bool gen2 = family == 5;
    // multi-rank: the separate pack kernel on the priority stream feeds the halo slots (no self-packing / interior-only variants)
    if (self_pack || A.clover || (ctx->nranks > 1 && !halo)) return LQCD_ERR_STATE;
    if (!g.regular || g.s[3] != 1 || g.c[3] != 1 || g.c[0] * g.c[1] * g.c[2] != TM_W || g.T < 2 || (g.part[3] && g.T < 3)) return LQCD_ERR_STATE;
    const int npatch = g.nt[0] * g.nt[1] * g.nt[2];
    // chunks per patch: minimise rounds x (steps + ~1 step of pipeline fill per task) over the divisors of T
    int nchunk = 1;
    {
        double best = 1e300;
        for (int c = 1; c <= g.T / 2; c++) {
            if (g.T % c) continue;
            const int slots = gen2 ? 2 * ctx->num_sms : ctx->num_sms;
            const int tasks = npatch * c, grid = tasks < slots ? tasks : slots;
            const double cost = (double)((tasks + grid - 1) / grid) * (g.T / c + (gen2 ? 0.5 : 1.0));
            if (cost < best) { best = cost; nchunk = c; }
        }
        if (const char *e = getenv("LQCD_TM_CHUNKS")) { int v = atoi(e); if (v >= 1 && g.T % v == 0) nchunk = v; }
    }
    TMArgs K;
    K.prefetch = 0;
    K.A = A; K.Lc = g.T / nchunk; K.nchunk = nchunk; K.nsb = g.nb[0] * g.nb[1] * g.nb[2]; K.ntasks = npatch * nchunk;
    if (gen2) {
        int g12 = 0;
        LQCD_TRY(ensure_links12(ctx, &g12));
        if (!g12) return LQCD_ERR_STATE;          // links are not SU(3): the register-resident kernel reads the full matrices
        K.A.links12 = ctx->links12;
        static int pf = -1;
        if (pf < 0) { const char *e = getenv("LQCD_TM_PREFETCH"); pf = (e && atoi(e) == 1) ? 1 : 0; }   // measured: 187.9 us with, 182.4 us without
        K.prefetch = pf;
        static int pipe = -1;
        if (pipe < 0) { const char *e = getenv("LQCD_TM_PIPE"); pipe = (e && atoi(e) == 1) ? 1 : 0; }
        static bool attr2_set = false;
        if (!attr2_set) {
#define TM2_ATTR(D, M, P) CUDA_TRY(ctx, cudaFuncSetAttribute(wilson_tmarch2_kernel<D, M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM2_SMEM_BYTES))
            TM2_ATTR(0, 0, 0); TM2_ATTR(1, 0, 0); TM2_ATTR(0, 1, 0); TM2_ATTR(1, 1, 0);
            TM2_ATTR(0, 0, 1); TM2_ATTR(1, 0, 1); TM2_ATTR(0, 1, 1); TM2_ATTR(1, 1, 1);
#undef TM2_ATTR
            attr2_set = true;
        }
        const int slots = 2 * ctx->num_sms, grid2 = K.ntasks < slots ? K.ntasks : slots;
#define TM2_GO(D, M, P) wilson_tmarch2_kernel<D, M, P><<<grid2, 128, TM2_SMEM_BYTES, s>>>(K)
        if (pipe) {
            if (halo) { if (dagger) TM2_GO(1, 1, 1); else TM2_GO(0, 1, 1); }
            else      { if (dagger) TM2_GO(1, 0, 1); else TM2_GO(0, 0, 1); }
        } else {
            if (halo) { if (dagger) TM2_GO(1, 1, 0); else TM2_GO(0, 1, 0); }
            else      { if (dagger) TM2_GO(1, 0, 0); else TM2_GO(0, 0, 0); }
        }
#undef TM2_GO
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
        return LQCD_OK;
    }
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(wilson_tmarch_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM_BYTES));
        CUDA_TRY(ctx, cudaFuncSetAttribute(wilson_tmarch_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM_BYTES));
        CUDA_TRY(ctx, cudaFuncSetAttribute(wilson_tmarch_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM_BYTES));
        CUDA_TRY(ctx, cudaFuncSetAttribute(wilson_tmarch_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM_BYTES));
        attr_set = true;
    }
    const int grid = K.ntasks < ctx->num_sms ? K.ntasks : ctx->num_sms;
    if (halo) {
        if (dagger) wilson_tmarch_kernel<1, 1><<<grid, 128, TM_SMEM_BYTES, s>>>(K);
        else        wilson_tmarch_kernel<0, 1><<<grid, 128, TM_SMEM_BYTES, s>>>(K);
    } else {
        if (dagger) wilson_tmarch_kernel<1, 0><<<grid, 128, TM_SMEM_BYTES, s>>>(K);
        else        wilson_tmarch_kernel<0, 0><<<grid, 128, TM_SMEM_BYTES, s>>>(K);
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
#ifdef TM_DEBUG
    {
        cudaError_t e = cudaStreamSynchronize(s);
        unsigned long long h[32];
        cudaMemcpyFromSymbol(h, tm_dbg, sizeof h);
        fprintf(stderr, "[tm_debug] grid %d Lc %d nchunk %d ntasks %d sync=%s timeouts=%llu", grid, K.Lc, K.nchunk, K.ntasks, cudaGetErrorString(e), h[0]);
        if (h[0]) {
            fprintf(stderr, " first: code %llu cta %llu thread %llu r %llu task %llu parity %llu bar %llu raw:", h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
            for (int j = 0; j < 7; j++) fprintf(stderr, " %016llx", h[8 + j]);
            unsigned long long z[32] = {0};
            cudaMemcpyToSymbol(tm_dbg, z, sizeof z);
        }
        fprintf(stderr, "\n");
    }
#endif
    return LQCD_OK;
}
