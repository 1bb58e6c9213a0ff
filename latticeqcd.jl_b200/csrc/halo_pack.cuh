// halo_pack.cuh -- device-side halo producer: spin-project the face spinors and STORE them straight into the
// neighbour ranks' halo slots over NVLink (peer-mapped pointers), then publish a sequence flag.
//
// Used in two ways:
//   * as the body of the first `npack` CTAs of the multi-GPU Dslash kernels (MULTI = 2, "self-packing"): ONE
//     kernel per operator application -- low blockIdx => dispatched first, so the halo is on the wire while the
//     interior tiles compute, and the face tiles (highest blockIdx) consume the neighbours' slots at the end;
//   * as the stand-alone halo_pack_kernel (comm.cu) on a high-priority stream (LQCD_SELF_PACK=0).
//
// Wilson: for the receiver's FORWARD hop we send P psi(m) from our LOW face (6 complex / site); for its
// BACKWARD hop we send U_mu^dag(m) P psi(m) from our HIGH face, so no remote links are ever needed.
// Staggered: 3 complex per site, same split.
#pragma once
#include "lqcd_internal.cuh"
#include "wilson_spin.cuh"

// face index -> site (coordinate mu fixed to cm); faces are enumerated lexicographically over the other three.
__device__ __forceinline__ int face_site(const Geom &g, int mu, int f, int cm) {
    const int d[4] = {g.X, g.Y, g.Z, g.T};
    int c[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i == mu) c[i] = cm;
        else { c[i] = f % d[i]; f /= d[i]; }
    }
    return c[0] + g.X * (c[1] + g.Y * (c[2] + g.Z * c[3]));
}

template <int MU>
__device__ __forceinline__ void wilson_pack_site(const cplx *__restrict__ in, const cplx *__restrict__ gauge, int dagger,
                                                 cplx *send, int side, int f, int s) {
    const cplx *sp = in + (size_t)(s >> 5) * (12 * 32) + (s & 31);
    cplx p[12];
#pragma unroll
    for (int k = 0; k < 12; k++) p[k] = ldg128(sp + k * 32);
    cplx o0[3], o1[3];
    if (side == 0) {                       // my low face: receiver's FORWARD hop, projector sign S = DAG ? +1 : -1
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (dagger) project<MU, +1>(o0[c], o1[c], p[c], p[3 + c], p[6 + c], p[9 + c]);
            else        project<MU, -1>(o0[c], o1[c], p[c], p[3 + c], p[6 + c], p[9 + c]);
        }
    } else {                               // my high face: receiver's BACKWARD hop: U^dag(m) P psi(m), S = DAG ? -1 : +1
        cplx h0[3], h1[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (dagger) project<MU, -1>(h0[c], h1[c], p[c], p[3 + c], p[6 + c], p[9 + c]);
            else        project<MU, +1>(h0[c], h1[c], p[c], p[3 + c], p[6 + c], p[9 + c]);
        }
        const cplx *lk = gauge + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cplx g0 = cmake(0, 0), g1 = cmake(0, 0);
#pragma unroll
            for (int b = 0; b < 3; b++) {
                cplx u = ldg128(lk + (b * 3 + a) * 32);
                cfmac(g0, u, h0[b]); cfmac(g1, u, h1[b]);
            }
            o0[a] = g0; o1[a] = g1;
        }
    }
    cplx *dst = send + (size_t)(f >> 5) * (6 * 32) + (f & 31);
#pragma unroll
    for (int c = 0; c < 3; c++) { dst[c * 32] = o0[c]; dst[(3 + c) * 32] = o1[c]; }
}

template <int MU>
__device__ __forceinline__ void stag_pack_site(const cplx *__restrict__ in, const cplx *__restrict__ gauge, cplx *send,
                                               int side, int f, int s) {
    const cplx *sp = in + (size_t)(s >> 5) * (3 * 32) + (s & 31);
    cplx v[3], o[3];
#pragma unroll
    for (int c = 0; c < 3; c++) v[c] = ldg128(sp + c * 32);
    if (side == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) o[c] = v[c];
    } else {
        const cplx *lk = gauge + ((size_t)(s >> 5) * 4 + MU) * (9 * 32) + (s & 31);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cplx acc = cmake(0, 0);
#pragma unroll
            for (int b = 0; b < 3; b++) cfmac(acc, ldg128(lk + (b * 3 + a) * 32), v[b]);
            o[a] = acc;
        }
    }
    cplx *dst = send + (size_t)(f >> 5) * (6 * 32) + (f & 31);
#pragma unroll
    for (int c = 0; c < 3; c++) dst[c * 32] = o[c];
}

// Work of pack CTA number `pcta` (0 <= pcta < H.cta0[4]) with blockDim.x threads; all threads of the CTA must call it.
__device__ __forceinline__ void halo_pack_cta(const Geom &g, int kind, int dagger, const cplx *__restrict__ in,
                                              const cplx *__restrict__ gauge, const HaloOut &H, int pcta) {
    const unsigned long long ts0 = (H.timing && threadIdx.x == 0) ? global_ns() : 0ull;
    int mu = 0;
    while (mu < 3 && pcta >= H.cta0[mu + 1]) mu++;
    const int d[4] = {g.X, g.Y, g.Z, g.T};
    const int F = g.V / d[mu];
    // H.spt face sites per thread: fewer, longer pack CTAs leave the remaining CTA slots of the first wave to the interior tiles
    for (int k = 0; k < H.spt; k++) {
        const int i = ((pcta - H.cta0[mu]) * H.spt + k) * blockDim.x + threadIdx.x;
        if (i >= 2 * F) break;
        const int side = i / F, f = i % F;
        const int s = face_site(g, mu, f, side ? d[mu] - 1 : 0);
        cplx *send = H.send[mu][side];
        if (kind == LQCD_WILSON) {
            switch (mu) {
            case 0: wilson_pack_site<0>(in, gauge, dagger, send, side, f, s); break;
            case 1: wilson_pack_site<1>(in, gauge, dagger, send, side, f, s); break;
            case 2: wilson_pack_site<2>(in, gauge, dagger, send, side, f, s); break;
            default: wilson_pack_site<3>(in, gauge, dagger, send, side, f, s); break;
            }
        } else {
            switch (mu) {
            case 0: stag_pack_site<0>(in, gauge, send, side, f, s); break;
            case 1: stag_pack_site<1>(in, gauge, send, side, f, s); break;
            case 2: stag_pack_site<2>(in, gauge, send, side, f, s); break;
            default: stag_pack_site<3>(in, gauge, send, side, f, s); break;
            }
        }
    }
    // publish: bar.sync orders the CTA's peer stores before thread 0's system-scope fence (cumulative), then the
    // ticket; the last pack CTA to arrive raises the sequence flags at the neighbours.
    __syncthreads();
    __shared__ int pack_last;
    if (threadIdx.x == 0) {
        // sys scope: wait here for this CTA's NVLink write acks.  gpu scope (H.gpu_fence): only order the stores before
        // the ticket; the last CTA observes every ticket and its system-scope fence below is cumulative over them.
        if (H.gpu_fence) __threadfence(); else __threadfence_system();
        const unsigned int npack = (unsigned int)H.cta0[4];
        pack_last = (atomicInc(H.ticket, npack - 1) == npack - 1);
    }
    __syncthreads();
    if (pack_last && threadIdx.x < 8) {
        const int m = threadIdx.x >> 1, side = threadIdx.x & 1;
        __threadfence_system();
        if (g.part[m]) st_release_sys(H.send_flag[m][side], H.seq);
    }
    if (H.timing && threadIdx.x == 0) stamp_span(H.timing, 0, ts0, global_ns());
}
