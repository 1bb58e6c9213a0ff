// wilson_spin.cuh -- (1 +- gamma_mu) spin projection / reconstruction in the upstream gamma basis
// (tables recalled from LatticeDiracOperators.jl's WilsonFermion constructor).
// Shared by the bulk Dslash kernel, the halo pack / exterior kernels and the force kernel.
#pragma once
#include "lqcd_internal.cuh"

// spin projection (1 + S*gamma_MU) psi -> two colour vectors, upstream gamma basis:
//   MU=0: h0 = p0 - S i p3, h1 = p1 - S i p2 ; rows 2,3 = ( S i h1,  S i h0)
//   MU=1: h0 = p0 - S p3,   h1 = p1 + S p2   ; rows 2,3 = ( S h1,   -S h0)
//   MU=2: h0 = p0 - S i p2, h1 = p1 + S i p3 ; rows 2,3 = ( S i h0, -S i h1)
//   MU=3: h0 = p0 - S p2,   h1 = p1 - S p3   ; rows 2,3 = (-S h0,   -S h1)
template <int MU, int S>
__device__ __forceinline__ void project(cplx &h0, cplx &h1, cplx p0, cplx p1, cplx p2, cplx p3) {
    if (MU == 0) {
        h0 = (S > 0) ? cadd(p0, cmulmi(p3)) : cadd(p0, cmuli(p3));
        h1 = (S > 0) ? cadd(p1, cmulmi(p2)) : cadd(p1, cmuli(p2));
    } else if (MU == 1) {
        h0 = (S > 0) ? csub(p0, p3) : cadd(p0, p3);
        h1 = (S > 0) ? cadd(p1, p2) : csub(p1, p2);
    } else if (MU == 2) {
        h0 = (S > 0) ? cadd(p0, cmulmi(p2)) : cadd(p0, cmuli(p2));
        h1 = (S > 0) ? cadd(p1, cmuli(p3)) : cadd(p1, cmulmi(p3));
    } else {
        h0 = (S > 0) ? csub(p0, p2) : cadd(p0, p2);
        h1 = (S > 0) ? csub(p1, p3) : cadd(p1, p3);
    }
}

template <int MU, int S>
__device__ __forceinline__ void reconstruct(cplx (&acc)[12], int a, cplx g0, cplx g1) {
    acc[0 + a] = cadd(acc[0 + a], g0);
    acc[3 + a] = cadd(acc[3 + a], g1);
    if (MU == 0) {        // rows 2,3 = ( S i g1, S i g0 )
        acc[6 + a] = (S > 0) ? cadd(acc[6 + a], cmuli(g1)) : cadd(acc[6 + a], cmulmi(g1));
        acc[9 + a] = (S > 0) ? cadd(acc[9 + a], cmuli(g0)) : cadd(acc[9 + a], cmulmi(g0));
    } else if (MU == 1) { // ( S g1, -S g0 )
        acc[6 + a] = (S > 0) ? cadd(acc[6 + a], g1) : csub(acc[6 + a], g1);
        acc[9 + a] = (S > 0) ? csub(acc[9 + a], g0) : cadd(acc[9 + a], g0);
    } else if (MU == 2) { // ( S i g0, -S i g1 )
        acc[6 + a] = (S > 0) ? cadd(acc[6 + a], cmuli(g0)) : cadd(acc[6 + a], cmulmi(g0));
        acc[9 + a] = (S > 0) ? cadd(acc[9 + a], cmulmi(g1)) : cadd(acc[9 + a], cmuli(g1));
    } else {              // ( -S g0, -S g1 )
        acc[6 + a] = (S > 0) ? csub(acc[6 + a], g0) : cadd(acc[6 + a], g0);
        acc[9 + a] = (S > 0) ? csub(acc[9 + a], g1) : cadd(acc[9 + a], g1);
    }
}

