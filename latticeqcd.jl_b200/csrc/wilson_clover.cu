// wilson_clover.cu -- Wilson-clover Dslash: the CLOVER = 1 instantiations of the Wilson kernel template.
//
//     y(n) = A(n) x(n) - kappa * sum_mu [ (1 - g_mu) U_mu(n) x(n+mu) + (1 + g_mu) U_mu^dag(n-mu) x(n-mu) ]
//
// A(n) = 1 + kappa*csw * sum_{mu<nu} sigma_mu_nu (x) i F^_mu_nu(n) is Hermitian and commutes with gamma_5, so the same
// epilogue serves D and D^dag.  New capability behind op.csw (BASELINE.json configs[3]; SURVEY.md 8a: the surveyed wrapper
// cannot reach Wilson-clover, parameter parsed at src/system/parameter_structs.jl:125); the term is built by clover.cu.
// Compulsory traffic 960 + 576 = 1536 B/site, 1368 + 504 flop/site.
//
// Parity: tests/test_gpu_extended.py (clover term, Dslash, CG against the oracle on hardware), tests/test_clover.py (oracle vs a numpy
// restatement of the same packing).  The convention is the textbook one: the reference cannot reach clover, so nothing upstream pins it.
#include "wilson_kernel.cuh"

int launch_wilson_clover(lqcd_ctx *ctx, const WilsonArgs &A, int dagger, int multi, int lh, int grid, int bs, cudaStream_t s) {
    if (bs > 256) return lqcd_fail(ctx, LQCD_ERR_ARG, "LQCD_WPC > 8 is not supported by the Wilson kernel");
#define CK(MT, MB, MU_, LH_)                                                                      \
    do {                                                                                          \
        if (dagger) wilson_dslash_kernel<1, MT, MB, MU_, LH_, 1><<<grid, bs, 0, s>>>(A);          \
        else        wilson_dslash_kernel<0, MT, MB, MU_, LH_, 1><<<grid, bs, 0, s>>>(A);          \
    } while (0)
#define CL(MT, MB)                                                                                \
    do {                                                                                          \
        if (lh) { if (multi == 2) CK(MT, MB, 2, 1); else if (multi == 1) CK(MT, MB, 1, 1); else CK(MT, MB, 0, 1); } \
        else    { if (multi == 2) CK(MT, MB, 2, 0); else if (multi == 1) CK(MT, MB, 1, 0); else CK(MT, MB, 0, 0); } \
    } while (0)
    if (bs > 128) CL(256, 1); else CL(128, 3);
#undef CL
#undef CK
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}
