// gauge_io.cu -- gauge configurations between the reference's file formats and the device layout.
//
// Reference behaviour being mirrored: `initial = "<file>"` + `loadU_format` in {"ILDG", "BridgeText", "JLD"} loads the start
// configuration (src/system/universe.jl:58-77: ILDG :62-65, load_BridgeText! :66-68) and `saveU_format` writes one every
// `saveU_every` trajectories (src/system/lqcd.jl:226-247: save_binarydata for "ILDG", save_textdata for "BridgeText"); all
// fermionic tests of the reference start from such a file (test/test_wilson.toml:17,27 -> test/confs_*/conf_00000100.ildg.txt).
// Formats as the reference's own fixtures show them:
//     both      site-major, x fastest, then y, z, t; inside a site mu = 1..4, row a, column b, (re, im)
//     ILDG      ONE LIME record: 144-byte header (magic 0x456789ab, version 1, flags MB|ME = 0xc000, 64-bit big-endian length,
//               type "ildg-binary-data" zero-padded to 128 bytes) + big-endian float64 payload (no ildg-format XML record)
//     Bridge++  text, one float64 per line, printed the way Julia prints a Float64 (shortest round-trip digits; positional
//               notation while the decimal point sits within (-4, 6], "d.ddde-5" otherwise)
// JLD2 (HDF5 container of Julia objects) is not handled here.
//
// Two levels:
//   lqcd_io_read_gauge / lqcd_io_write_gauge    HOST only, no GPU needed: file <-> the four Julia-layout arrays U[mu][a,b,x,y,z,t]
//                                               (wing ndw, any NC) -- what load_BridgeText! / save_binarydata do
//   lqcd_gauge_load / lqcd_gauge_save           file <-> DEVICE links: every rank reads / writes only the rows of its own block,
//                                               the raw payload goes through one pinned staging buffer, and byte swap + the
//                                               site-major -> AoSoA-32 transposition happen in one kernel (one CTA per 32-site
//                                               block, 18 KB contiguous in, 18 KB contiguous out, staged through shared memory)
#include "lqcd_internal.cuh"
#include <charconv>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define LQCD_IO_ILDG 0
#define LQCD_IO_BRIDGETEXT 1

// ---- LIME / number formatting (host) -------------------------------------------------------------------------------------------
static inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

struct LimeRecord { long data_off; uint64_t length; };

// finds the ildg-binary-data record of a LIME file
static int lime_find_payload(const lqcd_ctx *ctx, FILE *f, const char *path, LimeRecord *out) {
    long pos = 0;
    unsigned char h[144];
    for (;;) {
        if (fseek(f, pos, SEEK_SET) != 0 || fread(h, 1, 144, f) != 144)
            return lqcd_fail(ctx, LQCD_ERR_ARG, "%s: no ildg-binary-data record found", path);
        const uint32_t magic = ((uint32_t)h[0] << 24) | ((uint32_t)h[1] << 16) | ((uint32_t)h[2] << 8) | h[3];
        if (magic != 0x456789abu) return lqcd_fail(ctx, LQCD_ERR_ARG, "%s: not a LIME file (bad magic at offset %ld)", path, pos);
        uint64_t len = 0;
        for (int i = 0; i < 8; i++) len = (len << 8) | h[8 + i];
        char type[129];
        memcpy(type, h + 16, 128); type[128] = 0;
        if (strcmp(type, "ildg-binary-data") == 0) { out->data_off = pos + 144; out->length = len; return LQCD_OK; }
        pos += 144 + (long)((len + 7) / 8 * 8);
    }
}

static void lime_header(unsigned char h[144], uint64_t len) {
    memset(h, 0, 144);
    h[0] = 0x45; h[1] = 0x67; h[2] = 0x89; h[3] = 0xab;      // magic
    h[4] = 0x00; h[5] = 0x01;                                // version 1
    h[6] = 0xc0; h[7] = 0x00;                                // message begin | message end
    for (int i = 0; i < 8; i++) h[8 + i] = (unsigned char)(len >> (56 - 8 * i));
    memcpy(h + 16, "ildg-binary-data", 16);
}

// Julia's print(::Float64) (Base.Ryu.writeshortest with its defaults): shortest round-trip digits; positional notation when
// the decimal point position pt satisfies -4 < pt <= 6, else d.ddde<exp> without exponent padding; always a fractional part.
static int julia_float(char *dst, double v) {
    if (v != v) return sprintf(dst, "NaN");
    if (v == 1.0 / 0.0) return sprintf(dst, "Inf");
    if (v == -1.0 / 0.0) return sprintf(dst, "-Inf");
    char *p = dst;
    if (std::signbit(v)) { *p++ = '-'; v = -v; }
    if (v == 0.0) { memcpy(p, "0.0", 3); return (int)(p + 3 - dst); }
    char sci[40];
    auto r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);      // d[.ddd]e[+-]XX, shortest digits
    *r.ptr = 0;
    char digits[24] = {0};
    int nd = 0;
    const char *q = sci;
    for (; *q && *q != 'e'; q++)
        if (*q != '.') digits[nd++] = *q;
    const int e10 = atoi(q + 1);
    const int pt = e10 + 1;                                   // digits before the decimal point
    if (-4 < pt && pt <= 6) {
        if (pt <= 0) {
            *p++ = '0'; *p++ = '.';
            for (int i = 0; i < -pt; i++) *p++ = '0';
            memcpy(p, digits, nd); p += nd;
        } else if (pt >= nd) {
            memcpy(p, digits, nd); p += nd;
            for (int i = 0; i < pt - nd; i++) *p++ = '0';
            *p++ = '.'; *p++ = '0';
        } else {
            memcpy(p, digits, pt); p += pt;
            *p++ = '.';
            memcpy(p, digits + pt, nd - pt); p += nd - pt;
        }
    } else {
        *p++ = digits[0]; *p++ = '.';
        if (nd == 1) *p++ = '0';
        else { memcpy(p, digits + 1, nd - 1); p += nd - 1; }
        *p++ = 'e';
        p += sprintf(p, "%d", e10);
    }
    return (int)(p - dst);
}

// ---- block reader / writer: the rows of a sub-block, file order, native-endian doubles ----------------------------------------
struct Block { int gd[4], ld[4], o[4]; int nc; };
static inline size_t site_doubles(int nc) { return (size_t)4 * nc * nc * 2; }

// raw = true: keep the file's byte order (ILDG: big-endian; the device kernel swaps)
static int read_block(const lqcd_ctx *ctx, const char *path, int format, const Block &b, double *dst, bool raw) {
    const size_t sd = site_doubles(b.nc);
    const size_t gV = (size_t)b.gd[0] * b.gd[1] * b.gd[2] * b.gd[3];
    FILE *f = fopen(path, format == LQCD_IO_ILDG ? "rb" : "r");
    if (!f) return lqcd_fail(ctx, LQCD_ERR_ARG, "cannot open %s: %s", path, strerror(errno));
    int rc = LQCD_OK;
    if (format == LQCD_IO_ILDG) {
        LimeRecord rec = {0, 0};
        rc = lime_find_payload(ctx, f, path, &rec);
        if (rc == LQCD_OK && rec.length != gV * sd * 8)
            rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: payload of %llu bytes does not match a %dx%dx%dx%d NC=%d lattice (%zu bytes)", path,
                           (unsigned long long)rec.length, b.gd[0], b.gd[1], b.gd[2], b.gd[3], b.nc, gV * sd * 8);
        const size_t row = (size_t)b.ld[0] * sd;
        for (int t = 0; t < b.ld[3] && rc == LQCD_OK; t++)
            for (int z = 0; z < b.ld[2] && rc == LQCD_OK; z++)
                for (int y = 0; y < b.ld[1] && rc == LQCD_OK; y++) {
                    const size_t gsite = (size_t)b.o[0] + b.gd[0] * ((size_t)(y + b.o[1]) + b.gd[1] * ((size_t)(z + b.o[2]) + (size_t)b.gd[2] * (t + b.o[3])));
                    double *d = dst + ((size_t)y + b.ld[1] * ((size_t)z + (size_t)b.ld[2] * t)) * row;
                    if (fseek(f, rec.data_off + (long)(gsite * sd * 8), SEEK_SET) != 0 || fread(d, 8, row, f) != row)
                        rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: short read", path);
                    else if (!raw) {
                        uint64_t *u = reinterpret_cast<uint64_t *>(d);
                        for (size_t i = 0; i < row; i++) u[i] = bswap64(u[i]);
                    }
                }
    } else if (format == LQCD_IO_BRIDGETEXT) {
        // sequential text: every line is parsed, the ones of this block are kept
        std::vector<char> buf(1 << 20);
        setvbuf(f, buf.data(), _IOFBF, buf.size());
        char line[128];
        for (size_t gsite = 0; gsite < gV && rc == LQCD_OK; gsite++) {
            int c[4];
            size_t s = gsite;
            for (int i = 0; i < 4; i++) { c[i] = (int)(s % b.gd[i]) - b.o[i]; s /= b.gd[i]; }
            const bool mine = c[0] >= 0 && c[0] < b.ld[0] && c[1] >= 0 && c[1] < b.ld[1] && c[2] >= 0 && c[2] < b.ld[2] && c[3] >= 0 && c[3] < b.ld[3];
            double *d = mine ? dst + ((size_t)c[0] + b.ld[0] * ((size_t)c[1] + b.ld[1] * ((size_t)c[2] + (size_t)b.ld[2] * c[3]))) * sd : nullptr;
            for (size_t k = 0; k < sd; k++) {
                if (!fgets(line, sizeof line, f)) { rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: file ends after %zu of %zu numbers", path, gsite * sd + k, gV * sd); break; }
                if (!mine) continue;
                char *end = nullptr;
                d[k] = strtod(line, &end);
                if (end == line) { rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: line %zu is not a number", path, gsite * sd + k + 1); break; }
            }
        }
        if (rc == LQCD_OK && fgets(line, sizeof line, f) && strspn(line, " \t\r\n") != strlen(line))
            rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: more numbers than a %dx%dx%dx%d NC=%d lattice holds", path, b.gd[0], b.gd[1], b.gd[2], b.gd[3], b.nc);
    } else {
        rc = lqcd_fail(ctx, LQCD_ERR_ARG, "unknown gauge file format %d (0 = ILDG, 1 = BridgeText)", format);
    }
    fclose(f);
    return rc;
}

// ILDG: every rank writes its own rows at their offsets (`header` on the rank that owns the origin); text: whole lattice only
static int write_block(const lqcd_ctx *ctx, const char *path, int format, const Block &b, const double *src, bool raw, bool header) {
    const size_t sd = site_doubles(b.nc);
    const size_t gV = (size_t)b.gd[0] * b.gd[1] * b.gd[2] * b.gd[3];
    if (format == LQCD_IO_ILDG) {
        FILE *f = fopen(path, header ? "wb" : "r+b");
        if (!f) return lqcd_fail(ctx, LQCD_ERR_ARG, "cannot open %s for writing: %s", path, strerror(errno));
        int rc = LQCD_OK;
        if (header) {
            unsigned char h[144];
            lime_header(h, gV * sd * 8);
            if (fwrite(h, 1, 144, f) != 144) rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: write failed", path);
        }
        const size_t row = (size_t)b.ld[0] * sd;
        std::vector<uint64_t> tmp(raw ? 0 : row);
        for (int t = 0; t < b.ld[3] && rc == LQCD_OK; t++)
            for (int z = 0; z < b.ld[2] && rc == LQCD_OK; z++)
                for (int y = 0; y < b.ld[1] && rc == LQCD_OK; y++) {
                    const size_t gsite = (size_t)b.o[0] + b.gd[0] * ((size_t)(y + b.o[1]) + b.gd[1] * ((size_t)(z + b.o[2]) + (size_t)b.gd[2] * (t + b.o[3])));
                    const double *d = src + ((size_t)y + b.ld[1] * ((size_t)z + (size_t)b.ld[2] * t)) * row;
                    const void *w = d;
                    if (!raw) {
                        const uint64_t *u = reinterpret_cast<const uint64_t *>(d);
                        for (size_t i = 0; i < row; i++) tmp[i] = bswap64(u[i]);
                        w = tmp.data();
                    }
                    if (fseek(f, 144 + (long)(gsite * sd * 8), SEEK_SET) != 0 || fwrite(w, 8, row, f) != row)
                        rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: write failed", path);
                }
        if (fclose(f) != 0 && rc == LQCD_OK) rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: close failed: %s", path, strerror(errno));
        return rc;
    }
    if (format != LQCD_IO_BRIDGETEXT) return lqcd_fail(ctx, LQCD_ERR_ARG, "unknown gauge file format %d (0 = ILDG, 1 = BridgeText)", format);
    for (int i = 0; i < 4; i++)
        if (b.ld[i] != b.gd[i]) return lqcd_fail(ctx, LQCD_ERR_ARG, "BridgeText is sequential: save it from a single rank (or through lqcd_gauge_download)");
    FILE *f = fopen(path, "w");
    if (!f) return lqcd_fail(ctx, LQCD_ERR_ARG, "cannot open %s for writing: %s", path, strerror(errno));
    std::vector<char> buf(1 << 20);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    char num[48];
    int rc = LQCD_OK;
    for (size_t i = 0; i < gV * sd && rc == LQCD_OK; i++) {
        const int n = julia_float(num, src[i]);
        num[n] = '\n';
        if (fwrite(num, 1, n + 1, f) != (size_t)(n + 1)) rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: write failed", path);
    }
    if (fclose(f) != 0 && rc == LQCD_OK) rc = lqcd_fail(ctx, LQCD_ERR_ARG, "%s: close failed: %s", path, strerror(errno));
    return rc;
}

// ---- host arrays (Julia layout U[mu][a, b, x+w, y+w, z+w, t+w], column major) ------------------------------------------------
static int check_host_args(const char *path, const int dims[4], int nc, const void *U_mu, int ndw) {
    if (!path || !dims || !U_mu) return lqcd_fail(nullptr, LQCD_ERR_ARG, "null argument");
    if (nc < 1 || nc > 8) return lqcd_fail(nullptr, LQCD_ERR_ARG, "bad NC %d", nc);
    if (ndw < 0 || ndw > 4) return lqcd_fail(nullptr, LQCD_ERR_ARG, "bad wing width %d", ndw);
    for (int i = 0; i < 4; i++)
        if (dims[i] < 1) return lqcd_fail(nullptr, LQCD_ERR_ARG, "bad lattice extent %d", dims[i]);
    return LQCD_OK;
}
static inline size_t host_site(const int d[4], int w, int x, int y, int z, int t) {
    return (size_t)(x + w) + (size_t)(d[0] + 2 * w) * ((size_t)(y + w) + (size_t)(d[1] + 2 * w) * ((size_t)(z + w) + (size_t)(d[2] + 2 * w) * (t + w)));
}

extern "C" int lqcd_io_read_gauge(const char *path, int format, const int dims[4], int nc, double *const U_mu[4], int ndw) {
    LQCD_TRY(check_host_args(path, dims, nc, U_mu, ndw));
    Block b;
    for (int i = 0; i < 4; i++) { b.gd[i] = b.ld[i] = dims[i]; b.o[i] = 0; }
    b.nc = nc;
    const size_t V = (size_t)dims[0] * dims[1] * dims[2] * dims[3], sd = site_doubles(nc);
    std::vector<double> tmp(V * sd);
    LQCD_TRY(read_block(nullptr, path, format, b, tmp.data(), false));
    const int n2 = nc * nc;
    for (int mu = 0; mu < 4; mu++)
        if (!U_mu[mu]) return lqcd_fail(nullptr, LQCD_ERR_ARG, "U_mu[%d] is null", mu);
    size_t s = 0;
    for (int t = 0; t < dims[3]; t++) for (int z = 0; z < dims[2]; z++) for (int y = 0; y < dims[1]; y++) for (int x = 0; x < dims[0]; x++, s++) {
        const size_t hs = host_site(dims, ndw, x, y, z, t);
        for (int mu = 0; mu < 4; mu++) {
            const double *src = tmp.data() + (s * 4 + mu) * n2 * 2;
            double *dst = U_mu[mu] + hs * n2 * 2;
            for (int a = 0; a < nc; a++) for (int c = 0; c < nc; c++) {      // file: row a, column c;  Julia memory: a fastest
                dst[(a + nc * c) * 2] = src[(a * nc + c) * 2];
                dst[(a + nc * c) * 2 + 1] = src[(a * nc + c) * 2 + 1];
            }
        }
    }
    return LQCD_OK;
}

extern "C" int lqcd_io_write_gauge(const char *path, int format, const int dims[4], int nc, const double *const U_mu[4], int ndw) {
    LQCD_TRY(check_host_args(path, dims, nc, U_mu, ndw));
    Block b;
    for (int i = 0; i < 4; i++) { b.gd[i] = b.ld[i] = dims[i]; b.o[i] = 0; }
    b.nc = nc;
    const size_t V = (size_t)dims[0] * dims[1] * dims[2] * dims[3], sd = site_doubles(nc);
    std::vector<double> tmp(V * sd);
    const int n2 = nc * nc;
    for (int mu = 0; mu < 4; mu++)
        if (!U_mu[mu]) return lqcd_fail(nullptr, LQCD_ERR_ARG, "U_mu[%d] is null", mu);
    size_t s = 0;
    for (int t = 0; t < dims[3]; t++) for (int z = 0; z < dims[2]; z++) for (int y = 0; y < dims[1]; y++) for (int x = 0; x < dims[0]; x++, s++) {
        const size_t hs = host_site(dims, ndw, x, y, z, t);
        for (int mu = 0; mu < 4; mu++) {
            double *dst = tmp.data() + (s * 4 + mu) * n2 * 2;
            const double *src = U_mu[mu] + hs * n2 * 2;
            for (int a = 0; a < nc; a++) for (int c = 0; c < nc; c++) {
                dst[(a * nc + c) * 2] = src[(a + nc * c) * 2];
                dst[(a * nc + c) * 2 + 1] = src[(a + nc * c) * 2 + 1];
            }
        }
    }
    return write_block(nullptr, path, format, b, tmp.data(), false, true);
}

// ---- device path ---------------------------------------------------------------------------------------------------------------
// file order (local block, x fastest; per site mu, a, b) <-> AoSoA-32 links.  One CTA per 32-site block: its 32 x 36 complex
// numbers are one contiguous 18 KB piece on both sides.  SWAP: the staged doubles are big-endian (ILDG payload as read / to be
// written).
__device__ __forceinline__ double swap_double(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    unsigned int lo = (unsigned int)u, hi = (unsigned int)(u >> 32);
    lo = __byte_perm(lo, 0, 0x0123); hi = __byte_perm(hi, 0, 0x0123);
    return __longlong_as_double((long long)(((unsigned long long)lo << 32) | hi));
}

template <int TO_DEVICE, int SWAP>
__global__ void __launch_bounds__(256) gauge_file_kernel(cplx *__restrict__ dev, cplx *__restrict__ file, int nblk) {
    __shared__ cplx sm[32 * 36 + 32];            // +1 complex of padding per site row
    const int blk = blockIdx.x;
    if (blk >= nblk) return;
    cplx *fb = file + (size_t)blk * (32 * 36);
    cplx *db = dev + (size_t)blk * (36 * 32);
    if (TO_DEVICE) {
        for (int i = threadIdx.x; i < 32 * 36; i += blockDim.x) {
            cplx v = fb[i];
            if (SWAP) { v.x = swap_double(v.x); v.y = swap_double(v.y); }
            sm[i + i / 36] = v;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 36 * 32; i += blockDim.x) {
            const int k = i >> 5, lane = i & 31;          // k = mu*9 + 3a + b
            db[i] = sm[lane * 37 + k];
        }
    } else {
        for (int i = threadIdx.x; i < 36 * 32; i += blockDim.x) {
            const int k = i >> 5, lane = i & 31;
            sm[lane * 37 + k] = db[i];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 36; i += blockDim.x) {
            cplx v = sm[i + i / 36];
            if (SWAP) { v.x = swap_double(v.x); v.y = swap_double(v.y); }
            fb[i] = v;
        }
    }
}

static Block ctx_block(const lqcd_ctx *ctx) {
    Block b;
    const Geom &g = ctx->g;
    b.gd[0] = g.gX; b.gd[1] = g.gY; b.gd[2] = g.gZ; b.gd[3] = g.gT;
    b.ld[0] = g.X; b.ld[1] = g.Y; b.ld[2] = g.Z; b.ld[3] = g.T;
    for (int i = 0; i < 4; i++) b.o[i] = g.o[i];
    b.nc = 3;
    return b;
}

// load_gaugefield! / load_BridgeText! straight into the device links (universe.jl:62-68).  Collective across ranks only in the
// sense that every rank must call it; no data is exchanged.
extern "C" int lqcd_gauge_load(lqcd_ctx *ctx, const char *path, int format) {
    if (!ctx || !path) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const Block b = ctx_block(ctx);
    const size_t bytes = (size_t)ctx->g.V * 36 * sizeof(cplx);
    double *host = nullptr;
    CUDA_TRY(ctx, cudaMallocHost(&host, bytes));
    int rc = read_block(ctx, path, format, b, host, format == LQCD_IO_ILDG);
    cplx *stage = nullptr;
    if (rc == LQCD_OK && cudaMalloc(&stage, bytes) != cudaSuccess) rc = lqcd_fail(ctx, LQCD_ERR_CUDA, "cudaMalloc(%zu) failed", bytes);
    if (rc == LQCD_OK) {
        cudaError_t e = cudaMemcpyAsync(stage, host, bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            if (format == LQCD_IO_ILDG) gauge_file_kernel<1, 1><<<ctx->g.nblk, 256, 0, ctx->stream>>>(ctx->gauge, stage, ctx->g.nblk);
            else                        gauge_file_kernel<1, 0><<<ctx->g.nblk, 256, 0, ctx->stream>>>(ctx->gauge, stage, ctx->g.nblk);
            ctx->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = lqcd_fail(ctx, LQCD_ERR_CUDA, "gauge load: %s", cudaGetErrorString(e));
    }
    cudaFree(stage);
    cudaFreeHost(host);
    if (rc == LQCD_OK) { ctx->gauge_valid = true; ctx->gauge_epoch++; }
    return rc;
}

// save_binarydata / save_textdata from the device links (lqcd.jl:236-242).  ILDG: every rank writes its own rows into the same
// file; the rank that owns the lattice origin creates it, so callers place a host barrier between that rank's call and the
// others' (the Python / Julia mirrors do).  BridgeText: single rank.
extern "C" int lqcd_gauge_save(lqcd_ctx *ctx, const char *path, int format) {
    if (!ctx || !path) return lqcd_fail(ctx, LQCD_ERR_ARG, "null argument");
    if (!ctx->gauge_valid) return lqcd_fail(ctx, LQCD_ERR_STATE, "no gauge field on the device");
    if (format != LQCD_IO_ILDG && format != LQCD_IO_BRIDGETEXT) return lqcd_fail(ctx, LQCD_ERR_ARG, "unknown gauge file format %d (0 = ILDG, 1 = BridgeText)", format);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const Block b = ctx_block(ctx);
    const size_t bytes = (size_t)ctx->g.V * 36 * sizeof(cplx);
    double *host = nullptr;
    cplx *stage = nullptr;
    CUDA_TRY(ctx, cudaMallocHost(&host, bytes));
    int rc = LQCD_OK;
    if (cudaMalloc(&stage, bytes) != cudaSuccess) rc = lqcd_fail(ctx, LQCD_ERR_CUDA, "cudaMalloc(%zu) failed", bytes);
    if (rc == LQCD_OK) {
        if (format == LQCD_IO_ILDG) gauge_file_kernel<0, 1><<<ctx->g.nblk, 256, 0, ctx->stream>>>(ctx->gauge, stage, ctx->g.nblk);
        else                        gauge_file_kernel<0, 0><<<ctx->g.nblk, 256, 0, ctx->stream>>>(ctx->gauge, stage, ctx->g.nblk);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(host, stage, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = lqcd_fail(ctx, LQCD_ERR_CUDA, "gauge save: %s", cudaGetErrorString(e));
    }
    if (rc == LQCD_OK) {
        const bool origin = b.o[0] == 0 && b.o[1] == 0 && b.o[2] == 0 && b.o[3] == 0;
        rc = write_block(ctx, path, format, b, host, format == LQCD_IO_ILDG, origin);
    }
    cudaFree(stage);
    cudaFreeHost(host);
    return rc;
}
