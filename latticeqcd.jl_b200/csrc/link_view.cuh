// link_view.cuh -- the links of the whole process grid as seen from one rank: every rank's AoSoA-32 link array is peer
// mapped (CUDA IPC, comm.cu), so setup / gauge-sector kernels that reach one site across a face (clover leaves, staples,
// plaquettes) simply load the neighbour rank's links over NVLink.  The CALLER orders writers and readers (host barrier for
// the clover build, the in-kernel all-reduce between MD sub-steps).
#pragma once
#include "lqcd_internal.cuh"
#include <cstring>

struct LinkView {                                 // links of the whole process grid seen from this rank
    const cplx *base[LQCD_MAX_RANKS];             // base[r]: rank r's AoSoA-32 link array (peer mapped; base[my rank] = local)
    int pg[4], pc[4];                             // process grid and my coordinates in it
};

// 3x3 link at LOCAL coordinates that may be one step outside the local lattice in any direction
__device__ __forceinline__ void fetch_link(cplx (&m)[3][3], const LinkView &L, const Geom &g, const int (&c)[4], int mu) {
    const int d[4] = {g.X, g.Y, g.Z, g.T};
    int lc[4], rank = 0, mul = 1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int p = L.pc[i], v = c[i];
        if (v < 0) { v += d[i]; p = (p + L.pg[i] - 1) % L.pg[i]; }
        else if (v >= d[i]) { v -= d[i]; p = (p + 1) % L.pg[i]; }
        lc[i] = v; rank += p * mul; mul *= L.pg[i];
    }
    const int s = lc[0] + g.X * (lc[1] + g.Y * (lc[2] + g.Z * lc[3]));
    const cplx *p = L.base[rank] + ((size_t)(s >> 5) * 4 + mu) * (9 * 32) + (s & 31);
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) m[a][b] = p[(a * 3 + b) * 32];
}


int comm_link_view(lqcd_ctx *ctx, const cplx **bases);     // comm.cu: every rank's link array (peer mapped); error if unavailable

// fills a LinkView for this context (single rank: the local array only)
static inline int make_link_view(lqcd_ctx *ctx, LinkView &L) {
    memset(&L, 0, sizeof L);
    for (int i = 0; i < 4; i++) { L.pg[i] = ctx->procgrid[i]; L.pc[i] = ctx->pcoord[i]; }
    if (ctx->nranks > 1) return comm_link_view(ctx, L.base);
    L.base[0] = ctx->gauge;
    return LQCD_OK;
}
