// context.cu -- context, geometry/tiling, host<->device layout conversion, synthetic field generators.
//
// Boundary side of the library: replaces the storage part of Gaugefields.jl's Initialize_Gaugefields
// (src/system/universe.jl:41-49) and LatticeDiracOperators.jl's Initialize_pseudofermion_fields
// (universe.jl:107,112) with device-resident AoSoA-32 mirrors (layout: lqcd_internal.cuh).
#include "lqcd_internal.cuh"
#include "reduce.cuh"
#include "rng.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static std::string g_last_error;

int lqcd_fail(const lqcd_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_last_error = buf;
    return code;
}

extern "C" const char *lqcd_last_error(const lqcd_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_last_error.c_str();
}
extern "C" int lqcd_abi_version(void) { return LQCD_ABI_VERSION; }

// ---- tiling ------------------------------------------------------------------------------------------
void make_tiling(Geom &g) {
    const int d[4] = {g.X, g.Y, g.Z, g.T};
    int rem = 32;
    g.regular = 1;
    for (int i = 0; i < 4; i++) {
        int si = d[i] < rem ? d[i] : rem;
        if (rem % si != 0 || d[i] % si != 0) { g.regular = 0; break; }
        g.s[i] = si; rem /= si;
    }
    if (rem != 1) g.regular = 0;
    int wpc = 4;
    if (const char *e = getenv("LQCD_WPC")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) wpc = v; }
    g.wpc = wpc;
    for (int i = 0; i < 4; i++) { g.c[i] = 1; g.nt[i] = 1; g.nb[i] = 1; }
    if (!g.regular) { for (int i = 0; i < 4; i++) g.s[i] = 0; return; }
    for (int i = 0; i < 4; i++) g.nb[i] = d[i] / g.s[i];
    int tile[4] = {1, 1, 1, 1};
    bool user = false;
    if (const char *e = getenv("LQCD_TILE")) {
        int a, b, c, dd;
        if (sscanf(e, "%d,%d,%d,%d", &a, &b, &c, &dd) == 4 && a * b * c * dd == wpc && g.nb[0] % a == 0 &&
            g.nb[1] % b == 0 && g.nb[2] % c == 0 && g.nb[3] % dd == 0) {
            tile[0] = a; tile[1] = b; tile[2] = c; tile[3] = dd; user = true;
        }
    }
    if (!user) {
        int left = wpc, dir = 1, stuck = 0;
        while (left > 1 && stuck < 4) {     // spread factors of two over y, z, t, (x) round-robin
            if (g.nb[dir] % (tile[dir] * 2) == 0) { tile[dir] *= 2; left /= 2; stuck = 0; } else stuck++;
            dir = (dir + 1) % 4; if (dir == 0 && g.nb[0] == 1) dir = 1;
        }
        if (left > 1) { g.regular = 0; return; }
    }
    for (int i = 0; i < 4; i++) { g.c[i] = tile[i]; g.nt[i] = g.nb[i] / tile[i]; }
}

int reduce_grid(const lqcd_ctx *ctx) { return ctx->num_sms * 4; }

extern "C" int lqcd_ctx_create(const int gd[4], const int pg[4], int rank, int device, lqcd_ctx **out) {
    if (!gd || !pg || !out) return lqcd_fail(nullptr, LQCD_ERR_ARG, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return lqcd_fail(nullptr, LQCD_ERR_NOGPU, "no CUDA device (%s); liblqcd_b200 has no CPU fallback",
                         e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return lqcd_fail(nullptr, LQCD_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    int nranks = pg[0] * pg[1] * pg[2] * pg[3];
    if (nranks < 1 || rank < 0 || rank >= nranks) return lqcd_fail(nullptr, LQCD_ERR_ARG, "bad rank/procgrid");
    for (int i = 0; i < 4; i++)
        if (gd[i] < 2 || pg[i] < 1 || gd[i] % pg[i] != 0 || (pg[i] > 1 && gd[i] / pg[i] < 2))
            return lqcd_fail(nullptr, LQCD_ERR_ARG, "dims[%d]=%d not divisible by procgrid %d (local extent must be >= 2)", i, gd[i], pg[i]);
    lqcd_ctx *ctx = new lqcd_ctx();
    ctx->device = device; ctx->rank = rank; ctx->nranks = nranks;
    int r = rank;
    for (int i = 0; i < 4; i++) { ctx->procgrid[i] = pg[i]; ctx->pcoord[i] = r % pg[i]; r /= pg[i]; }
    Geom &g = ctx->g;
    g.gX = gd[0]; g.gY = gd[1]; g.gZ = gd[2]; g.gT = gd[3];
    g.X = gd[0] / pg[0]; g.Y = gd[1] / pg[1]; g.Z = gd[2] / pg[2]; g.T = gd[3] / pg[3];
    const int loc[4] = {g.X, g.Y, g.Z, g.T};
    for (int i = 0; i < 4; i++) { g.part[i] = pg[i] > 1; g.o[i] = ctx->pcoord[i] * loc[i]; }
    long long V = 1LL * g.X * g.Y * g.Z * g.T;
    if (V % 32 != 0 || V > (1LL << 30)) { delete ctx; return lqcd_fail(nullptr, LQCD_ERR_ARG, "local volume %lld must be a multiple of 32 and < 2^30", V); }
    g.V = (int)V; g.nblk = g.V / 32;
    make_tiling(g);
    ctx->gauge = nullptr; ctx->gauge_valid = false; ctx->stage = nullptr; ctx->stage_bytes = 0;
    ctx->flush = nullptr; ctx->flush_bytes = 0; ctx->launches = 0; ctx->comm = nullptr; ctx->force_buf = nullptr; ctx->force_valid = false; ctx->mom = nullptr; ctx->mom_valid = false;
    ctx->eo = nullptr; ctx->eo_active = 0; ctx->stag_even_solve = 0; ctx->pipe = nullptr; ctx->queue = nullptr; ctx->mrhs = nullptr;
    ctx->gauge_epoch = 0; ctx->clover = nullptr; ctx->clover_k = nullptr; ctx->clover_epoch = ~0ull; ctx->clover_coef = 0.0;
    ctx->hist_dev = nullptr; ctx->hist_cap = 0;
    ctx->links12 = nullptr; ctx->links12_epoch = ~0ull; ctx->links12_ok = false; ctx->links12_dev = 0.0; ctx->links12_scratch = nullptr;
    // a failure half way through creation goes through the normal destructor (null-safe frees), so streams / events / allocations
    // made so far are released
#define CT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { int rc = lqcd_fail(nullptr, LQCD_ERR_CUDA, "%s -> %s", #expr, cudaGetErrorString(_e)); lqcd_ctx_destroy(ctx); return rc; } } while (0)
    CT(cudaSetDevice(device));
    cudaDeviceProp prop;
    CT(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { lqcd_ctx_destroy(ctx); return lqcd_fail(nullptr, LQCD_ERR_NOGPU, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor); }
    ctx->num_sms = prop.multiProcessorCount;
    CT(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {   // stream2 carries the halo pack kernels: highest priority so that its CTAs are dispatched before the
        // Dslash kernel's remaining tiles (the face tiles of the Dslash kernel wait for the NEIGHBOUR's pack;
        // if the local pack were starved behind them two GPUs could wait for each other).
        int lo = 0, hi = 0;
        CT(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CT(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi));
    }
    CT(cudaEventCreate(&ctx->ev0)); CT(cudaEventCreate(&ctx->ev1));
    CT(cudaEventCreateWithFlags(&ctx->ev_pack, cudaEventDisableTiming));
    CT(cudaEventCreateWithFlags(&ctx->ev_int, cudaEventDisableTiming));
    CT(cudaEventCreateWithFlags(&ctx->ev_poll[0], cudaEventDisableTiming));
    CT(cudaEventCreateWithFlags(&ctx->ev_poll[1], cudaEventDisableTiming));
    CT(cudaMalloc(&ctx->gauge, (size_t)g.nblk * 4 * 9 * 32 * sizeof(cplx)));
    // reduction workspace: partials sized for the largest grid any kernel uses
    size_t maxgrid = (size_t)g.nblk + 1024;
    CT(cudaMalloc(&ctx->red.partials, maxgrid * LQCD_MAX_RED * sizeof(double)));
    CT(cudaMalloc(&ctx->red.ticket, sizeof(unsigned int)));
    CT(cudaMemset(ctx->red.ticket, 0, sizeof(unsigned int)));
    CT(cudaMalloc(&ctx->queue, sizeof(unsigned int)));
    CT(cudaMemset(ctx->queue, 0, sizeof(unsigned int)));
    CT(cudaMalloc(&ctx->red.st, sizeof(SolverState)));
    CT(cudaMemset(ctx->red.st, 0, sizeof(SolverState)));
    ctx->red.hist = nullptr;
    memset(&ctx->red.cr, 0, sizeof ctx->red.cr);
    CT(cudaMallocHost(&ctx->st_host, 3 * sizeof(SolverState)));
#undef CT
    *out = ctx;
    return LQCD_OK;
}

int comm_destroy(lqcd_ctx *ctx);
void eo_destroy(lqcd_ctx *ctx);       // wilson_eo.cu
void pipe_destroy(lqcd_ctx *ctx);     // host_pipeline.cu
void mrhs_destroy(lqcd_ctx *ctx);     // mrhs.cu

extern "C" int lqcd_ctx_destroy(lqcd_ctx *ctx) {
    if (!ctx) return LQCD_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    comm_destroy(ctx);
    eo_destroy(ctx);
    pipe_destroy(ctx);
    mrhs_destroy(ctx);
    for (int k = 0; k < 2; k++)
        for (auto *f : ctx->scratch[k]) if (f) { cudaFree(f->d); delete f; }
    cudaFree(ctx->links12); cudaFree(ctx->links12_scratch);
    cudaFree(ctx->gauge); cudaFree(ctx->stage); cudaFree(ctx->flush); cudaFree(ctx->hist_dev); cudaFree(ctx->force_buf); cudaFree(ctx->mom); cudaFree(ctx->clover); cudaFree(ctx->clover_k);
    cudaFree(ctx->queue); cudaFree(ctx->red.partials); cudaFree(ctx->red.ticket); cudaFree(ctx->red.st);
    cudaFreeHost(ctx->st_host);
    cudaEvent_t evs[6] = {ctx->ev0, ctx->ev1, ctx->ev_pack, ctx->ev_int, ctx->ev_poll[0], ctx->ev_poll[1]};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);                    // (null after a creation that failed half way)
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    cudaGetLastError();
    delete ctx;
    return LQCD_OK;
}

extern "C" int lqcd_local_dims(const lqcd_ctx *ctx, int ld[4], int origin[4]) {
    if (!ctx) return lqcd_fail(nullptr, LQCD_ERR_ARG, "null ctx");
    ld[0] = ctx->g.X; ld[1] = ctx->g.Y; ld[2] = ctx->g.Z; ld[3] = ctx->g.T;
    if (origin) for (int i = 0; i < 4; i++) origin[i] = ctx->g.o[i];
    return LQCD_OK;
}

extern "C" int lqcd_synchronize(lqcd_ctx *ctx) {
    if (!ctx) return lqcd_fail(nullptr, LQCD_ERR_ARG, "null ctx");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream2));
    return LQCD_OK;
}

extern "C" int lqcd_host_register(lqcd_ctx *ctx, void *ptr, size_t bytes) {
    if (!ctx || !ptr) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return LQCD_OK;
}
extern "C" int lqcd_host_unregister(lqcd_ctx *ctx, void *ptr) {
    if (!ctx || !ptr) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    CUDA_TRY(ctx, cudaHostUnregister(ptr));
    return LQCD_OK;
}
extern "C" int lqcd_launch_count(const lqcd_ctx *ctx, uint64_t *count) {
    if (!ctx || !count) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    *count = ctx->launches;
    return LQCD_OK;
}
extern "C" int lqcd_stream(lqcd_ctx *ctx, void **s) {
    if (!ctx || !s) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    *s = (void *)ctx->stream;
    return LQCD_OK;
}

static int ensure_stage(lqcd_ctx *ctx, size_t bytes) {
    if (ctx->stage_bytes >= bytes) return LQCD_OK;
    if (ctx->stage) CUDA_TRY(ctx, cudaFree(ctx->stage));
    ctx->stage = nullptr; ctx->stage_bytes = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->stage, bytes));
    ctx->stage_bytes = bytes;
    return LQCD_OK;
}

// ---- layout conversion kernels --------------------------------------------------------------------
// host site index with wing w
__device__ __forceinline__ size_t host_site(const Geom &g, int s, int w) {
    int x = s % g.X; s /= g.X;
    int y = s % g.Y; s /= g.Y;
    int z = s % g.Z; int t = s / g.Z;
    return (size_t)(x + w) + (size_t)(g.X + 2 * w) * ((y + w) + (size_t)(g.Y + 2 * w) * ((z + w) + (size_t)(g.Z + 2 * w) * (t + w)));
}

// links: host [a + 3*(b + 3*hsite)] (one array per mu, staged back to back) <-> device AoSoA-32.
// One warp per (block, mu): the 288 complex of the block are read/written coalesced on the host-layout side.
template <int TO_DEVICE>
__global__ void convert_links_kernel(cplx *dev, cplx *host, Geom g, int w, size_t host_stride) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp_global >= g.nblk * 4) return;
    const int blk = warp_global >> 2, mu = warp_global & 3;
    cplx *d = dev + ((size_t)blk * 4 + mu) * (9 * 32);
    cplx *h = host + (size_t)mu * host_stride;
    if (w == 0) {
        cplx *hb = h + (size_t)blk * 32 * 9;
        for (int j = lane; j < 288; j += 32) {
            int sl = j / 9, eh = j % 9, a = eh % 3, b = eh / 3;
            if (TO_DEVICE) d[(a * 3 + b) * 32 + sl] = hb[j]; else hb[j] = d[(a * 3 + b) * 32 + sl];
        }
    } else {
        size_t hs = host_site(g, blk * 32 + lane, w);
        for (int eh = 0; eh < 9; eh++) {
            int a = eh % 3, b = eh / 3;
            if (TO_DEVICE) d[(a * 3 + b) * 32 + lane] = h[hs * 9 + eh]; else h[hs * 9 + eh] = d[(a * 3 + b) * 32 + lane];
        }
    }
}

// fermions: host [c + 3*(hsite + Vh*alpha)] <-> device [(blk*ncomp + 3*alpha + c)*32 + lane]
template <int TO_DEVICE>
__global__ void convert_fermion_kernel(cplx *dev, cplx *host, Geom g, int w, int nspin, size_t Vh) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp_global >= g.nblk * nspin) return;
    const int blk = warp_global / nspin, al = warp_global % nspin;
    cplx *d = dev + ((size_t)blk * nspin * 3 + al * 3) * 32;
    if (w == 0) {
        cplx *hb = host + ((size_t)blk * 32 + Vh * al) * 3;
        for (int j = lane; j < 96; j += 32) {
            int sl = j / 3, c = j % 3;
            if (TO_DEVICE) d[c * 32 + sl] = hb[j]; else hb[j] = d[c * 32 + sl];
        }
    } else {
        size_t hs = host_site(g, blk * 32 + lane, w);
        for (int c = 0; c < 3; c++) {
            size_t hi = c + 3 * (hs + Vh * al);
            if (TO_DEVICE) d[c * 32 + lane] = host[hi]; else host[hi] = d[c * 32 + lane];
        }
    }
}

int convert_fermion_range(lqcd_ctx *ctx, int to_device, cplx *dev, cplx *host_stage, int ncomp, int blk0, int nblk, cudaStream_t s) {
    const int nspin = ncomp / 3;
    Geom g = ctx->g;
    g.nblk = nblk;                                                  // the kernel covers blocks [0, nblk) of the shifted pointers
    cplx *d = dev + (size_t)blk0 * ncomp * 32, *h = host_stage + (size_t)blk0 * 32 * 3;
    const int warps = nblk * nspin, bs = 256, grid = (warps * 32 + bs - 1) / bs;
    if (to_device) convert_fermion_kernel<1><<<grid, bs, 0, s>>>(d, h, g, 0, nspin, (size_t)ctx->g.V);
    else           convert_fermion_kernel<0><<<grid, bs, 0, s>>>(d, h, g, 0, nspin, (size_t)ctx->g.V);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return LQCD_OK;
}

static size_t host_volume(const Geom &g, int w) {
    return (size_t)(g.X + 2 * w) * (g.Y + 2 * w) * (g.Z + 2 * w) * (g.T + 2 * w);
}

// the four host arrays (Julia layout, wing ndw) -> a device link-layout field
int upload_links_to(lqcd_ctx *ctx, cplx *dev_links, const double *const U_mu[4], int ndw) {
    if (ndw < 0 || ndw > 4) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad wing width %d", ndw);
    const size_t Vh = host_volume(ctx->g, ndw), per = Vh * 9 * sizeof(cplx);
    LQCD_TRY(ensure_stage(ctx, per * 4));
    for (int mu = 0; mu < 4; mu++) {
        if (!U_mu[mu]) return lqcd_fail(ctx, LQCD_ERR_ARG, "U_mu[%d] is null", mu);
        CUDA_TRY(ctx, cudaMemcpyAsync((char *)ctx->stage + mu * per, U_mu[mu], per, cudaMemcpyHostToDevice, ctx->stream));
    }
    int warps = ctx->g.nblk * 4, bs = 256, grid = (warps * 32 + bs - 1) / bs;
    convert_links_kernel<1><<<grid, bs, 0, ctx->stream>>>(dev_links, (cplx *)ctx->stage, ctx->g, ndw, Vh * 9);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

extern "C" int lqcd_gauge_upload(lqcd_ctx *ctx, const double *const U_mu[4], int nc, int ndw) {
    if (!ctx || !U_mu) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (nc != 3) return lqcd_fail(ctx, LQCD_ERR_ARG, "only NC = 3 is implemented on the device path (got %d)", nc);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LQCD_TRY(upload_links_to(ctx, ctx->gauge, U_mu, ndw));
    ctx->gauge_valid = true; ctx->gauge_epoch++;
    return LQCD_OK;
}

// device link-layout field -> the four host arrays (Julia layout, wing ndw)
static int links_to_host(lqcd_ctx *ctx, const cplx *dev, double *const U_mu[4], int ndw) {
    const size_t Vh = host_volume(ctx->g, ndw), per = Vh * 9 * sizeof(cplx);
    LQCD_TRY(ensure_stage(ctx, per * 4));
    if (ndw > 0) CUDA_TRY(ctx, cudaMemsetAsync(ctx->stage, 0, per * 4, ctx->stream));
    int warps = ctx->g.nblk * 4, bs = 256, grid = (warps * 32 + bs - 1) / bs;
    convert_links_kernel<0><<<grid, bs, 0, ctx->stream>>>(const_cast<cplx *>(dev), (cplx *)ctx->stage, ctx->g, ndw, Vh * 9);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    for (int mu = 0; mu < 4; mu++) {
        if (!U_mu[mu]) return lqcd_fail(ctx, LQCD_ERR_ARG, "output array %d is null", mu);
        CUDA_TRY(ctx, cudaMemcpyAsync(U_mu[mu], (char *)ctx->stage + mu * per, per, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

int download_links_from(lqcd_ctx *ctx, const cplx *dev_links, double *const U_mu[4], int ndw) { return links_to_host(ctx, dev_links, U_mu, ndw); }

extern "C" int lqcd_gauge_download(lqcd_ctx *ctx, double *const U_mu[4], int nc, int ndw) {
    if (!ctx || !U_mu) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (nc != 3) return lqcd_fail(ctx, LQCD_ERR_ARG, "only NC = 3");
    if (ndw < 0 || ndw > 4) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad wing width %d", ndw);
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "no gauge field on the device");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return links_to_host(ctx, ctx->gauge, U_mu, ndw);
}

// ---- fermion fields ----------------------------------------------------------------------------------
static int check_f(const lqcd_ctx *ctx, const lqcd_fermion *f) {
    if (!ctx || !f) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (f->owner != ctx) return lqcd_fail(ctx, LQCD_ERR_ARG, "fermion field belongs to another context");
    return LQCD_OK;
}

static int alloc_fermion(lqcd_ctx *ctx, int kind, lqcd_fermion **out) {
    lqcd_fermion *f = new lqcd_fermion();
    f->kind = kind; f->ncomp = ncomp_of(kind); f->owner = ctx;
    f->bytes = (size_t)ctx->g.nblk * f->ncomp * 32 * sizeof(cplx);
    cudaError_t e = cudaMalloc(&f->d, f->bytes);
    if (e != cudaSuccess) {
        const size_t want = f->bytes;
        delete f;
        return lqcd_fail(ctx, LQCD_ERR_CUDA, "cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
    }
    *out = f;
    return LQCD_OK;
}

extern "C" int lqcd_fermion_alloc(lqcd_ctx *ctx, int kind, lqcd_fermion **out) {
    if (!ctx || !out) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (kind != LQCD_WILSON && kind != LQCD_STAGGERED) return lqcd_fail(ctx, LQCD_ERR_ARG, "unknown fermion kind %d", kind);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LQCD_TRY(alloc_fermion(ctx, kind, out));
    CUDA_TRY(ctx, cudaMemsetAsync((*out)->d, 0, (*out)->bytes, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

extern "C" int lqcd_fermion_free(lqcd_ctx *ctx, lqcd_fermion *f) {
    if (!f) return LQCD_OK;
    LQCD_TRY(check_f(ctx, f));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaFree(f->d));
    delete f;
    return LQCD_OK;
}

int get_scratch(lqcd_ctx *ctx, int kind, int idx, lqcd_fermion **out) {
    auto &v = ctx->scratch[kind];
    if ((int)v.size() <= idx) v.resize(idx + 1, nullptr);      // slots are allocated one by one on first use (callers use sparse ranges)
    if (!v[idx]) LQCD_TRY(alloc_fermion(ctx, kind, &v[idx]));
    *out = v[idx];
    return LQCD_OK;
}

extern "C" int lqcd_fermion_upload(lqcd_ctx *ctx, lqcd_fermion *f, const double *host, int ndw) {
    LQCD_TRY(check_f(ctx, f));
    if (!host || ndw < 0 || ndw > 4) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad host pointer / wing");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int nspin = f->ncomp / 3;
    const size_t Vh = host_volume(ctx->g, ndw), bytes = Vh * f->ncomp * sizeof(cplx);
    LQCD_TRY(ensure_stage(ctx, bytes));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int warps = ctx->g.nblk * nspin, bs = 256, grid = (warps * 32 + bs - 1) / bs;
    convert_fermion_kernel<1><<<grid, bs, 0, ctx->stream>>>(f->d, (cplx *)ctx->stage, ctx->g, ndw, nspin, Vh);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

extern "C" int lqcd_fermion_download(lqcd_ctx *ctx, const lqcd_fermion *f, double *host, int ndw) {
    LQCD_TRY(check_f(ctx, f));
    if (!host || ndw < 0 || ndw > 4) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad host pointer / wing");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int nspin = f->ncomp / 3;
    const size_t Vh = host_volume(ctx->g, ndw), bytes = Vh * f->ncomp * sizeof(cplx);
    LQCD_TRY(ensure_stage(ctx, bytes));
    if (ndw > 0) CUDA_TRY(ctx, cudaMemsetAsync(ctx->stage, 0, bytes, ctx->stream));
    int warps = ctx->g.nblk * nspin, bs = 256, grid = (warps * 32 + bs - 1) / bs;
    convert_fermion_kernel<0><<<grid, bs, 0, ctx->stream>>>(f->d, (cplx *)ctx->stage, ctx->g, ndw, nspin, Vh);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(host, ctx->stage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

extern "C" int lqcd_fermion_zero(lqcd_ctx *ctx, lqcd_fermion *f) {
    LQCD_TRY(check_f(ctx, f));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemsetAsync(f->d, 0, f->bytes, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

extern "C" int lqcd_fermion_copy(lqcd_ctx *ctx, lqcd_fermion *dst, const lqcd_fermion *src) {
    LQCD_TRY(check_f(ctx, dst)); LQCD_TRY(check_f(ctx, src));
    if (dst->kind != src->kind) return lqcd_fail(ctx, LQCD_ERR_ARG, "kind mismatch");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpyAsync(dst->d, src->d, src->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

// ---- counter-based generators (identical fields for any process grid) ------------------------------
__global__ void gauge_random_kernel(cplx *gauge, Geom g, uint64_t seed, double warm_eps) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;    // over V*4, site fastest within mu-major blocks
    if (idx >= g.V * 4) return;
    const int s = idx % g.V, mu = idx / g.V;
    const uint64_t link = (uint64_t)global_site(g, s) * 4 + mu;
    cplx m[3][3];
    for (int e = 0; e < 9; e++) {
        double a, b;
        gauss_pair(seed, link * 9 + e, a, b);
        m[e / 3][e % 3] = make_double2(a, b);
    }
    if (warm_eps >= 0.0) {     // 1 + i eps H, H hermitian traceless from the Gaussian matrix
        cplx h[3][3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                h[i][j] = make_double2(0.5 * (m[i][j].x + m[j][i].x), 0.5 * (m[i][j].y - m[j][i].y));
        double tr = (h[0][0].x + h[1][1].x + h[2][2].x) / 3.0;
        for (int i = 0; i < 3; i++) h[i][i].x -= tr;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                m[i][j] = make_double2((i == j ? 1.0 : 0.0) - warm_eps * h[i][j].y, warm_eps * h[i][j].x);
    }
    // Gram-Schmidt rows 0,1 ; row 2 = conj(row0 x row1)  -> SU(3)
    double n0 = 0;
    for (int j = 0; j < 3; j++) n0 += m[0][j].x * m[0][j].x + m[0][j].y * m[0][j].y;
    n0 = 1.0 / sqrt(n0);
    for (int j = 0; j < 3; j++) m[0][j] = cscale(n0, m[0][j]);
    cplx pr = make_double2(0, 0);
    for (int j = 0; j < 3; j++) cfmac(pr, m[0][j], m[1][j]);      // <row0,row1>
    for (int j = 0; j < 3; j++) m[1][j] = csub(m[1][j], cmul(pr, m[0][j]));
    double n1 = 0;
    for (int j = 0; j < 3; j++) n1 += m[1][j].x * m[1][j].x + m[1][j].y * m[1][j].y;
    n1 = 1.0 / sqrt(n1);
    for (int j = 0; j < 3; j++) m[1][j] = cscale(n1, m[1][j]);
    for (int j = 0; j < 3; j++) {
        int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        cplx c = csub(cmul(m[0][j1], m[1][j2]), cmul(m[0][j2], m[1][j1]));
        m[2][j] = make_double2(c.x, -c.y);
    }
    cplx *d = gauge + ((size_t)(s >> 5) * 4 + mu) * (9 * 32) + (s & 31);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) d[(a * 3 + b) * 32] = m[a][b];
}

extern "C" int lqcd_gauge_random(lqcd_ctx *ctx, uint64_t seed, double warm_eps) {
    if (!ctx) return lqcd_fail(nullptr, LQCD_ERR_ARG, "null ctx");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int n = ctx->g.V * 4, bs = 128;
    gauge_random_kernel<<<(n + bs - 1) / bs, bs, 0, ctx->stream>>>(ctx->gauge, ctx->g, seed, warm_eps);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->gauge_valid = true; ctx->gauge_epoch++;
    return LQCD_OK;
}

__global__ void fermion_gaussian_kernel(cplx *f, Geom g, int ncomp, uint64_t seed) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.V * ncomp) return;
    const int lane = idx & 31, k = (idx >> 5) % ncomp, blk = (idx >> 5) / ncomp;
    const uint64_t gs = (uint64_t)global_site(g, blk * 32 + lane);
    double a, b;
    gauss_pair(seed, gs * ncomp + k, a, b);
    const double sg = 0.70710678118654752440;   // sigma^2 = 1/2 per real component
    f[idx] = make_double2(sg * a, sg * b);
}

extern "C" int lqcd_fermion_gaussian(lqcd_ctx *ctx, lqcd_fermion *f, uint64_t seed) {
    LQCD_TRY(check_f(ctx, f));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int n = ctx->g.V * f->ncomp, bs = 256;
    fermion_gaussian_kernel<<<(n + bs - 1) / bs, bs, 0, ctx->stream>>>(f->d, ctx->g, f->ncomp, seed);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

// Z4_distribution_fermi!(x) (measure_chiral_condensate.jl:181): every component one of 1, i, -1, -i with equal probability --
// the noise sources of the stochastic trace.  Counter-based like the Gaussian field: the same on any process grid.
__global__ void fermion_z4_kernel(cplx *f, Geom g, int ncomp, uint64_t seed) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.V * ncomp) return;
    const int lane = idx & 31, k = (idx >> 5) % ncomp, blk = (idx >> 5) / ncomp;
    const uint64_t gs = (uint64_t)global_site(g, blk * 32 + lane);
    const unsigned q = (unsigned)(splitmix64(seed ^ splitmix64(gs * ncomp + k)) >> 62);      // top two bits
    f[idx] = make_double2(q == 0 ? 1.0 : (q == 2 ? -1.0 : 0.0), q == 1 ? 1.0 : (q == 3 ? -1.0 : 0.0));
}

extern "C" int lqcd_fermion_z4(lqcd_ctx *ctx, lqcd_fermion *f, uint64_t seed) {
    LQCD_TRY(check_f(ctx, f));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int n = ctx->g.V * f->ncomp, bs = 256;
    fermion_z4_kernel<<<(n + bs - 1) / bs, bs, 0, ctx->stream>>>(f->d, ctx->g, f->ncomp, seed);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

__global__ void fermion_mask_parity_kernel(cplx *f, Geom g, int ncomp, int parity) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // same indexing as fermion_gaussian_kernel
    if (idx >= g.V * ncomp) return;
    const int lane = idx & 31, blk = (idx >> 5) / ncomp;
    int s = blk * 32 + lane;
    const int x = s % g.X; s /= g.X;
    const int y = s % g.Y; s /= g.Y;
    const int z = s % g.Z; const int t = s / g.Z;
    const int p = (x + g.o[0] + y + g.o[1] + z + g.o[2] + t + g.o[3]) & 1;
    if (p != parity) f[idx] = make_double2(0.0, 0.0);
}

extern "C" int lqcd_fermion_mask_parity(lqcd_ctx *ctx, lqcd_fermion *f, int parity) {
    LQCD_TRY(check_f(ctx, f));
    if (parity != 0 && parity != 1) return lqcd_fail(ctx, LQCD_ERR_ARG, "parity must be 0 (even) or 1 (odd)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int n = ctx->g.V * f->ncomp, bs = 256;
    fermion_mask_parity_kernel<<<(n + bs - 1) / bs, bs, 0, ctx->stream>>>(f->d, ctx->g, f->ncomp, parity);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

extern "C" int lqcd_fermion_point_source(lqcd_ctx *ctx, lqcd_fermion *f, const int site[4], int color, int spin) {
    LQCD_TRY(check_f(ctx, f));
    const Geom &g = ctx->g;
    const int gd[4] = {g.gX, g.gY, g.gZ, g.gT}, ld[4] = {g.X, g.Y, g.Z, g.T};
    const int nspin = f->ncomp / 3;
    if (!site || color < 0 || color > 2 || spin < 0 || spin >= nspin) return lqcd_fail(ctx, LQCD_ERR_ARG, "bad source index");
    for (int i = 0; i < 4; i++) if (site[i] < 0 || site[i] >= gd[i]) return lqcd_fail(ctx, LQCD_ERR_ARG, "source site out of range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemsetAsync(f->d, 0, f->bytes, ctx->stream));
    bool mine = true;
    int l[4];
    for (int i = 0; i < 4; i++) { l[i] = site[i] - g.o[i]; if (l[i] < 0 || l[i] >= ld[i]) mine = false; }
    if (mine) {
        int s = l[0] + g.X * (l[1] + g.Y * (l[2] + g.Z * l[3]));
        cplx one = make_double2(1.0, 0.0);
        size_t off = ((size_t)(s >> 5) * f->ncomp + spin * 3 + color) * 32 + (s & 31);
        CUDA_TRY(ctx, cudaMemcpyAsync(f->d + off, &one, sizeof one, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LQCD_OK;
}

// ---- plaquette (pins loader + device link layout against SURVEY.md section 4 values) ----------------
__device__ __forceinline__ void load_link(cplx (&m)[3][3], const cplx *gauge, int s, int mu) {
    const cplx *d = gauge + ((size_t)(s >> 5) * 4 + mu) * (9 * 32) + (s & 31);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) m[a][b] = d[(a * 3 + b) * 32];
}
__global__ void plaquette_kernel(const cplx *gauge, Geom g, Reduce R) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double red[1] = {0.0};
    if (s < g.V) {
        int c[4], d[4] = {g.X, g.Y, g.Z, g.T}, st[4] = {1, g.X, g.X * g.Y, g.X * g.Y * g.Z};
        int r = s;
        for (int i = 0; i < 4; i++) { c[i] = r % d[i]; r /= d[i]; }
        for (int mu = 0; mu < 4; mu++)
            for (int nu = mu + 1; nu < 4; nu++) {
                int smu = (c[mu] == d[mu] - 1) ? s - (d[mu] - 1) * st[mu] : s + st[mu];
                int snu = (c[nu] == d[nu] - 1) ? s - (d[nu] - 1) * st[nu] : s + st[nu];
                cplx A[3][3], B[3][3], Cm[3][3], Dm[3][3], AB[3][3], ABC[3][3];
                load_link(A, gauge, s, mu); load_link(B, gauge, smu, nu);
                load_link(Cm, gauge, snu, mu); load_link(Dm, gauge, s, nu);
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
                    cplx acc = make_double2(0, 0);
                    for (int k = 0; k < 3; k++) cfma(acc, A[i][k], B[k][j]);
                    AB[i][j] = acc;
                }
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {      // AB * C^dag
                    cplx acc = make_double2(0, 0);
                    for (int k = 0; k < 3; k++) { cplx cc = make_double2(Cm[j][k].x, -Cm[j][k].y); cfma(acc, AB[i][k], cc); }
                    ABC[i][j] = acc;
                }
                for (int i = 0; i < 3; i++) {                                   // Re tr (ABC * D^dag)
                    for (int k = 0; k < 3; k++) red[0] += ABC[i][k].x * Dm[i][k].x + ABC[i][k].y * Dm[i][k].y;
                }
            }
    }
    grid_reduce_finish<1>(red, R, FIN_STORE);
}

extern "C" int lqcd_gauge_plaquette(lqcd_ctx *ctx, double *plaq) {
    if (!ctx || !plaq) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "no gauge field on the device");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->nranks > 1) {       // across ranks: the MD plaquette kernel reads the neighbours' peer-mapped links (gauge_md.cu); collective
        double act = 0.0;
        LQCD_TRY(lqcd_md_gauge_action(ctx, 1.0, &act));               // = -(1/NC) sum_plaq Re tr U_p over the GLOBAL lattice
        *plaq = -act / (6.0 * (double)ctx->g.gX * ctx->g.gY * ctx->g.gZ * ctx->g.gT);
        return LQCD_OK;
    }
    int bs = 128, grid = (ctx->g.V + bs - 1) / bs;
    plaquette_kernel<<<grid, bs, 0, ctx->stream>>>(ctx->gauge, ctx->g, ctx->red);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st_host, ctx->red.st, sizeof(SolverState), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *plaq = ctx->st_host->red[0] / (6.0 * 3.0 * (double)ctx->g.V);
    return LQCD_OK;
}
