// site_map.cuh -- CTA/warp -> 32-site block mapping and site coordinates.
//
// A CTA of `wpc` warps owns a small 4-d TILE of the block lattice (c[0..3] blocks per direction) so that
// the +-y/z/t neighbours of most of its sites are read by the same SM (L1 reuse); tiles are numbered with
// t slowest so the set of CTAs in flight sweeps the lattice in t and the neighbour slices stay in the
// 126 MB L2 (each spinor/link record is then fetched from HBM once per application).
#pragma once
#include "lqcd_internal.cuh"

__host__ __device__ __forceinline__ int block_of_warp(const Geom &g, int cta, int warp) {
    if (!g.regular) return cta * g.wpc + warp;
    int t0 = cta % g.nt[0]; cta /= g.nt[0];
    int t1 = cta % g.nt[1]; cta /= g.nt[1];
    int t2 = cta % g.nt[2];
    int t3 = cta / g.nt[2];
    int w0 = warp % g.c[0]; warp /= g.c[0];
    int w1 = warp % g.c[1]; warp /= g.c[1];
    int w2 = warp % g.c[2];
    int w3 = warp / g.c[2];
    int b0 = t0 * g.c[0] + w0, b1 = t1 * g.c[1] + w1, b2 = t2 * g.c[2] + w2, b3 = t3 * g.c[3] + w3;
    return b0 + g.nb[0] * (b1 + g.nb[1] * (b2 + g.nb[2] * b3));
}

__device__ __forceinline__ void site_coords(const Geom &g, int s, int &x, int &y, int &z, int &t) {
    x = s % g.X; s /= g.X;
    y = s % g.Y; s /= g.Y;
    z = s % g.Z;
    t = s / g.Z;
}

// face index of a site for direction MU: lexicographic over the other three coordinates.
template <int MU>
__host__ __device__ __forceinline__ int face_index(const Geom &g, int x, int y, int z, int t) {
    if (MU == 0) return y + g.Y * (z + g.Z * t);
    if (MU == 1) return x + g.X * (z + g.Z * t);
    if (MU == 2) return x + g.X * (y + g.Y * t);
    return x + g.X * (y + g.Y * z);
}

// face CTAs wait here for all neighbours' halo flags of this application (thread 0 spins, then bar.sync)
__device__ __forceinline__ void wait_halo_flags(const Geom &g, const HaloIn &H) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        const unsigned long long w0 = H.timing ? global_ns() : 0ull;
        bool good = true;
        for (int m = 0; m < 4 && good; m++) {
            if (!g.part[m]) continue;
            for (int side = 0; side < 2 && good; side++)
                while (ld_acquire_sys(H.recv_flag[m][side]) < H.seq)
                    if (clock64() - t0 > H.timeout_cycles) { good = false; *H.err = 1000000 + (m * 2 + side) * 100000 + (int)(H.seq % 100000); break; }
        }
        if (H.timing) { const unsigned long long dt = global_ns() - w0; atomicAdd(H.timing + 6, dt); atomicMax(H.timing + 7, dt); }
    }
    __syncthreads();
}
