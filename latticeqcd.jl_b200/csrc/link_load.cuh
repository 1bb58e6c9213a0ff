// link_load.cuh -- link loads of the Dslash kernels: cache-policy variants and the two-row ("links12") format.
#pragma once
#include "lqcd_internal.cuh"

// Cache policy of the link loads.  Links have at most one reuse (as the backward link of the +mu neighbour), spinors up
// to nine.  Measured on B200 (tools/quick_bench.py): marking link lines evict-first in L1 helps when the local lattice is
// L2 resident (32.32.16.8: 34.4 -> 31.3 us) and hurts at 32^4 (203 -> 228 us: the backward-link L1 hits are lost and L2
// is already the bottleneck); L1::no_allocate / spinor evict_last variants were slower in both regimes.  So LH = 1 is
// selected only for local volumes <= 2^18 sites (the strong-scaling regime).
template <int LH>
__device__ __forceinline__ cplx ldlink(const cplx *p) {
    if (LH == 1) {
        cplx v;
        asm("ld.global.nc.L1::evict_first.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
        return v;
    }
    return __ldg(p);
}
// third row of an SU(3) matrix from the first two: u[6 + b] = conj(row0 x row1)[b].  ONE definition: every kernel that reads two-row
// links (global or staged in shared memory) must produce the same bits.
__device__ __forceinline__ void link_row3(cplx (&u)[9]) {
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const cplx a0 = u[(b + 1) % 3], b1 = u[3 + (b + 2) % 3], c0 = u[(b + 2) % 3], d1 = u[3 + (b + 1) % 3];
        double re = a0.x * b1.x, im = a0.x * b1.y;                 // a0*b1 - c0*d1, conjugated
        re = fma(-a0.y, b1.y, re); im = fma(a0.y, b1.x, im);
        re = fma(-c0.x, d1.x, re); im = fma(-c0.x, d1.y, im);
        re = fma(c0.y, d1.y, re);  im = fma(-c0.y, d1.x, im);
        u[6 + b] = cmake(re, -im);
    }
}
// link of site `ls`, direction MU, as a row-major 3x3 in registers.  G12 = 1: `gauge` is the two-row copy (links12.cu, 6 loads),
// the third row is rebuilt as conj(row0 x row1) -- exact for SU(3) up to rounding, which is what ensure_links12 verified.
template <int MU, int G12, int LH>
__device__ __forceinline__ void load_link(cplx (&u)[9], const cplx *__restrict__ gauge, int ls) {
    if (G12) {
        const cplx *lk = gauge + ((size_t)(ls >> 5) * 4 + MU) * (6 * 32) + (ls & 31);
#pragma unroll
        for (int e = 0; e < 6; e++) u[e] = ldlink<LH>(lk + e * 32);
        link_row3(u);
    } else {
        const cplx *lk = gauge + ((size_t)(ls >> 5) * 4 + MU) * (9 * 32) + (ls & 31);
#pragma unroll
        for (int e = 0; e < 9; e++) u[e] = ldlink<LH>(lk + e * 32);
    }
}

