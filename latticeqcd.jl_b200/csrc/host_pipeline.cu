// host_pipeline.cu -- mul!(y, D, x) for HOST-resident fields as ONE pipelined operation.
//
// The reference-facing call of the Julia shim (LQCDB200.jl: mul!(y::AbstractFermionfields, D::B200Dirac, x)) has to move
// the source to the device and the result back (SURVEY.md 8b "Selection" option 1: CPU pseudofermion types, copy in/out).
// Done as upload -> Dslash -> download it costs H2D + D2H back to back, and at 32^4 the two 201 MB copies (~3.7 ms each
// at PCIe 5 x16 rates) are 40x the kernel time.  B200 has independent copy engines per direction and PCIe is full duplex,
// so here the lattice is cut into S slabs of t-slices and the three stages run as a pipeline on three streams:
//
//     copy-in stream   H2D slab 0 | H2D slab 1 | H2D slab 2 | ...
//     compute stream                convert 0  | convert 1  | convert 2, Dslash slab 1, convert-out 1 | ...
//     copy-out stream                                                     D2H slab 1 | ...
//
// The Dslash of slab k needs the source on slabs k-1, k, k+1 (periodic), so slab k is multiplied as soon as slab k+1 has
// been converted; slab 0 (whose lower neighbour is the LAST slab) goes last.  The Dslash kernels run on a sub-range of
// their t-slowest CTA order (DslashFuse.cta_off / cta_count), i.e. exactly the code path of the full-lattice launch.
// Host layout (Julia): Wilson psi[c, x, y, z, t, alpha] -> for every spin alpha a slab of t-slices is ONE contiguous piece.
// D^dag D: the first application needs the whole source, so only the copy-in of the first and the copy-out of the second
// application are overlapped with compute.
//
// Single rank, wing 0 (Wilson fields are "nowing", universe.jl:112), regular tiling; anything else takes the plain
// three-call sequence inside the same entry point, so callers never need to care.
#include "lqcd_internal.cuh"
#include <cstring>

struct HostPipe {
    cudaStream_t s_in, s_out;
    std::vector<cudaEvent_t> ev_in, ev_out;
    cplx *stage_in, *stage_out;
    size_t stage_bytes;
};

static int pipe_state(lqcd_ctx *ctx, HostPipe **out, int nslab, size_t bytes) {
    HostPipe *p = ctx->pipe;
    if (!p) {
        p = new HostPipe();
        p->stage_in = p->stage_out = nullptr; p->stage_bytes = 0;
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
        ctx->pipe = p;
    }
    while ((int)p->ev_in.size() < nslab) {
        cudaEvent_t a, b;
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        p->ev_in.push_back(a); p->ev_out.push_back(b);
    }
    if (p->stage_bytes < bytes) {
        if (p->stage_in) CUDA_TRY(ctx, cudaFree(p->stage_in));
        if (p->stage_out) CUDA_TRY(ctx, cudaFree(p->stage_out));
        p->stage_in = p->stage_out = nullptr; p->stage_bytes = 0;
        CUDA_TRY(ctx, cudaMalloc(&p->stage_in, bytes));
        CUDA_TRY(ctx, cudaMalloc(&p->stage_out, bytes));
        p->stage_bytes = bytes;
    }
    *out = p;
    return LQCD_OK;
}

void pipe_destroy(lqcd_ctx *ctx) {
    HostPipe *p = ctx->pipe;
    if (!p) return;
    cudaFree(p->stage_in); cudaFree(p->stage_out);
    for (auto e : p->ev_in) cudaEventDestroy(e);
    for (auto e : p->ev_out) cudaEventDestroy(e);
    cudaStreamDestroy(p->s_in); cudaStreamDestroy(p->s_out);
    delete p;
    ctx->pipe = nullptr;
}

static int slab_dslash(lqcd_ctx *ctx, const lqcd_op *op, cplx *y, const cplx *x, int dagger, int cta0, int ncta) {
    DslashFuse f = DslashFuse();
    f.cta_off = cta0; f.cta_count = ncta;
    if (op->kind == LQCD_WILSON) return launch_wilson_dslash(ctx, op, y, x, dagger, &f, ctx->stream);
    return launch_staggered_dslash(ctx, op, y, x, dagger, &f, ctx->stream);
}

extern "C" int lqcd_dslash_host(lqcd_ctx *ctx, const lqcd_op *op, lqcd_fermion *y, lqcd_fermion *x, double *y_host, const double *x_host,
                                int mode, int ndw) {
    if (!ctx || !op || !y || !x || !y_host || !x_host) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (y == x || (const double *)y_host == x_host) return lqcd_fail(ctx, LQCD_ERR_ARG, "mul!: output aliases input");
    if (mode < 0 || mode > 2) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad mode %d", mode);
    const Geom &g = ctx->g;
    // slabs of whole t-tiles.  The last three slabs (S-2, S-1 and 0, which wait for the final piece of the source) drain
    // after the copy-in has finished, so more slabs = shorter tail; 16 keeps every copy piece > 1 MB at 32^4.
    int S = 1;
    if (g.regular) for (int c = 16; c >= 2; c--) if (g.nt[3] % c == 0) { S = c; break; }
    const bool plain = ctx->nranks > 1 || ndw != 0 || !g.regular || S < 2 || (op->kind == LQCD_WILSON && op->r != 1.0) ||
                       y->kind != op->kind || x->kind != op->kind;
    if (plain) {                      // same result through the three-call sequence (every argument check happens there)
        LQCD_TRY(lqcd_fermion_upload(ctx, x, x_host, ndw));
        LQCD_TRY(lqcd_dslash(ctx, op, y, x, mode));
        return lqcd_fermion_download(ctx, y, y_host, ndw);
    }
    if (y->owner != ctx || x->owner != ctx) return lqcd_fail(ctx, LQCD_ERR_ARG, "field belongs to another context");
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "operator applied before lqcd_gauge_upload");
    for (int i = 0; i < 4; i++)
        if (op->bc[i] != 1.0 && op->bc[i] != -1.0) return lqcd_fail(ctx, LQCD_ERR_ARG, "boundary phase bc[%d] = %g must be +-1", i, op->bc[i]);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    HostPipe *P = nullptr;
    LQCD_TRY(pipe_state(ctx, &P, S, x->bytes));
    const int ncomp = x->ncomp, nspin = ncomp / 3;
    const int ncta_all = (g.nblk + g.wpc - 1) / g.wpc, cta_per = ncta_all / S;     // S divides nt[3], tiles are t-slowest
    const int blk_per = g.nblk / S;
    const size_t site_per = (size_t)blk_per * 32;
    const cplx *hx = (const cplx *)x_host;
    cplx *hy = (cplx *)y_host;
    lqcd_fermion *tmp = nullptr;
    if (mode == LQCD_OP_DDAGD) LQCD_TRY(get_scratch(ctx, op->kind, 0, &tmp));
    const int dag_last = (mode == LQCD_OP_D) ? 0 : 1;

    auto copy_in = [&](int k) -> int {        // host -> staged host-layout field, one contiguous piece per spin
        for (int al = 0; al < nspin; al++) {
            const size_t off = ((size_t)g.V * al + site_per * k) * 3;
            CUDA_TRY(ctx, cudaMemcpyAsync(P->stage_in + off, hx + off, site_per * 3 * sizeof(cplx), cudaMemcpyHostToDevice, P->s_in));
        }
        CUDA_TRY(ctx, cudaEventRecord(P->ev_in[k], P->s_in));
        return LQCD_OK;
    };
    auto emit = [&](int k, const cplx *src) -> int {     // y slab k = op(src) ; convert to host layout ; D2H on the copy-out stream
        LQCD_TRY(slab_dslash(ctx, op, y->d, src, dag_last, k * cta_per, cta_per));
        LQCD_TRY(convert_fermion_range(ctx, 0, y->d, P->stage_out, ncomp, k * blk_per, blk_per, ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(P->ev_out[k], ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(P->s_out, P->ev_out[k], 0));
        for (int al = 0; al < nspin; al++) {
            const size_t off = ((size_t)g.V * al + site_per * k) * 3;
            CUDA_TRY(ctx, cudaMemcpyAsync(hy + off, P->stage_out + off, site_per * 3 * sizeof(cplx), cudaMemcpyDeviceToHost, P->s_out));
        }
        return LQCD_OK;
    };

    for (int k = 0; k < S; k++) LQCD_TRY(copy_in(k));
    for (int k = 0; k < S; k++) {
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, P->ev_in[k], 0));
        LQCD_TRY(convert_fermion_range(ctx, 1, x->d, P->stage_in, ncomp, k * blk_per, blk_per, ctx->stream));
        if (mode != LQCD_OP_DDAGD && k >= 2) LQCD_TRY(emit(k - 1, x->d));     // slabs k-2, k-1, k are on the device
    }
    if (mode != LQCD_OP_DDAGD) {
        LQCD_TRY(emit(S - 1, x->d));                                          // needs slab 0: present
        LQCD_TRY(emit(0, x->d));                                              // needs slab S-1: present
    } else {
        if (op->kind == LQCD_WILSON) LQCD_TRY(launch_wilson_dslash(ctx, op, tmp->d, x->d, 0, nullptr, ctx->stream));
        else                         LQCD_TRY(launch_staggered_dslash(ctx, op, tmp->d, x->d, 0, nullptr, ctx->stream));
        for (int k = 0; k < S; k++) LQCD_TRY(emit(k, tmp->d));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(P->s_out));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(P->s_in));
    return LQCD_OK;
}
