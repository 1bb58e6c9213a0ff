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
    if (bs > 128) return lqcd_fail(ctx, LQCD_ERR_ARG, "LQCD_WPC > 4 is not supported by the Wilson kernel");
    // two-row links as in the plain Wilson launcher (wilson_dslash.cu)
    int g12 = 0;
    LQCD_TRY(ensure_links12(ctx, &g12));
    WilsonArgs B = A;
    B.links12 = g12 ? ctx->links12 : nullptr;
#define CK(MU_, LH_, G_)                                                                          \
    do {                                                                                          \
        if (dagger) wilson_dslash_kernel<1, MU_, LH_, 1, G_><<<grid, bs, 0, s>>>(B);              \
        else        wilson_dslash_kernel<0, MU_, LH_, 1, G_><<<grid, bs, 0, s>>>(B);              \
    } while (0)
#define CG(MU_, LH_) do { if (g12) CK(MU_, LH_, 1); else CK(MU_, LH_, 0); } while (0)
    if (lh) { if (multi == 2) CG(2, 1); else if (multi == 1) CG(1, 1); else CG(0, 1); }
    else    { if (multi == 2) CG(2, 0); else if (multi == 1) CG(1, 0); else CG(0, 0); }
#undef CG
#undef CK
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}
