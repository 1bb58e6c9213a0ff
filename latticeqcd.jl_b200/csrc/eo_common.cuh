// eo_common.cuh -- checkerboard (even / odd) half-lattice state shared by wilson_eo.cu and staggered_eo.cu.
#pragma once
#include "lqcd_internal.cuh"

struct EoState {
    Geom gh;                 // half-lattice geometry (X -> X/2), CTA tiling of its own
    int fullX;
    cplx *gauge[2];          // links owned by even / odd sites
    uint64_t epoch;          // gauge epoch the split links belong to
    cplx *f[5];              // half fields: 0 b_e, 1 b_o, 2 x_e, 3 t (hop temporary, odd), 4 bhat_e / x_o
    size_t nhalf;            // complex numbers per Wilson half field
    const cplx *stag_in;     // staggered_eo.cu: input of the last first-half application (needed by the second half)
};

int eo_state(lqcd_ctx *ctx, EoState **out);                                                              // wilson_eo.cu
int eo_convert(lqcd_ctx *ctx, EoState *e, int to_half, cplx *full, cplx *he, cplx *ho, int ncomp);        // wilson_eo.cu
int stag_even_apply(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, const DslashFuse *fuse);   // staggered_eo.cu
