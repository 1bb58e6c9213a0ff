// rng.cuh -- counter-based generators: the same (seed, global counter) gives the same number on any process grid, so
// synthetic fields are identical for every decomposition.
#pragma once
#include "lqcd_internal.cuh"

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ void gauss_pair(uint64_t seed, uint64_t ctr, double &g0, double &g1) {
    uint64_t a = splitmix64(seed ^ splitmix64(2 * ctr)), b = splitmix64(seed ^ splitmix64(2 * ctr + 1));
    double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0);      // (0,1]
    double u2 = (b >> 11) * (1.0 / 9007199254740992.0);              // [0,1)
    double rad = sqrt(-2.0 * log(u1)), s, c;
    sincospi(2.0 * u2, &s, &c);
    g0 = rad * c; g1 = rad * s;
}
__device__ __forceinline__ int global_site(const Geom &g, int s) {
    int x = s % g.X; s /= g.X;
    int y = s % g.Y; s /= g.Y;
    int z = s % g.Z; int t = s / g.Z;
    return (x + g.o[0]) + g.gX * ((y + g.o[1]) + g.gY * ((z + g.o[2]) + g.gZ * (t + g.o[3])));
}

