// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A minimal SIMT emulator + fake CUDA runtime that lets the library's .cu
// sources be compiled with g++ and executed on the CPU of the GPU-less build container, so that device code written
// while no B200 is reachable can still be checked against the oracle before it first runs on hardware (index algebra,
// reductions, barriers, finish ops, host-side solver drivers, memory bounds via red zones / ASan).
//
// It is NOT a backend: nothing in latticeqcd.jl_b200/ includes or links it, the product library is built by nvcc only and
// refuses to run without an sm_100 device; `-m gpu` tests, bench.py and smoke() never load the emulated library.  The
// emulation executes CTAs one after the other in blockIdx order (threads of a CTA as cooperatively scheduled fibers),
// i.e. ONE legal schedule of the CUDA execution model: it validates functional correctness, not memory-model races,
// performance or PTX semantics of the inline-asm helpers (those are replaced by tests/emu/translate.py, see there).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define LQCD_EMU 1

struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- runtime types -------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801,
       cudaErrorPeerAccessAlreadyEnabled = 704, cudaErrorLaunchFailure = 719 };
struct emu_stream; struct emu_event;
typedef emu_stream *cudaStream_t;
typedef emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostRegisterDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major, minor, multiProcessorCount, clockRate; size_t sharedMemPerBlockOptin; char name[64]; };
struct cudaIpcMemHandle_t { char reserved[64]; };

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t emu_malloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t emu_malloc_host(void **p, size_t bytes);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaHostRegister(void *p, size_t bytes, unsigned flags);
cudaError_t cudaHostUnregister(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void *p, int v, size_t n);
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s = nullptr);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return emu_malloc((void **)p, bytes); }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return emu_malloc_host((void **)p, bytes); }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- launch + SIMT intrinsics -------------------------------------------------------------------------------------------
namespace emu {
void launch(dim3 grid, dim3 block, size_t dyn_smem, cudaStream_t s, const std::function<void()> &thread_body, const char *name = "?");
void cta_barrier();
void warp_barrier();
void yield_now();
uint64_t *warp_slot(int lane);          // exchange buffer of the calling fiber's warp
unsigned char *dynamic_smem();
int lane_id();
}   // namespace emu

static inline void __syncthreads() { emu::cta_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __threadfence() { __sync_synchronize(); }
static inline void __threadfence_system() { __sync_synchronize(); }
static inline void __threadfence_block() { __sync_synchronize(); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *(const volatile T *)p; }
static inline double2 __ldcg(const double2 *p) { const volatile double *q = (const volatile double *)p; return make_double2(q[0], q[1]); }
template <class T> static inline T __ldcs(const T *p) { return *p; }
static inline long long clock64() { return (long long)__builtin_ia32_rdtsc(); }
static inline unsigned atomicInc(unsigned *p, unsigned wrap) {
    unsigned old = *p;                      // fibers are cooperative and CTAs sequential inside a process
    *p = (old >= wrap) ? 0u : old + 1u;
    return old;
}
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
static inline void sincospi(double x, double *s, double *c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }

template <class T> static inline T __shfl_xor_sync(unsigned, T v, int mask) {
    static_assert(sizeof(T) <= 8, "emulated shuffles move at most 8 bytes");
    const int lane = emu::lane_id();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    *emu::warp_slot(lane) = raw;
    emu::warp_barrier();
    raw = *emu::warp_slot(lane ^ mask);
    emu::warp_barrier();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
    const int lane = emu::lane_id();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    *emu::warp_slot(lane) = raw;
    emu::warp_barrier();
    raw = *emu::warp_slot(src & 31);
    emu::warp_barrier();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    const int lane = emu::lane_id();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    *emu::warp_slot(lane) = raw;
    emu::warp_barrier();
    raw = *emu::warp_slot(lane + d < 32 ? lane + d : lane);
    emu::warp_barrier();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}

// ---- mbarrier / bulk-copy model (wilson_dslash3.cu helpers are mapped onto these by translate.py) -------------------------
namespace emu {
struct MBar { uint32_t expected, arrived, tx_pending, phase; };     // lives in the kernel's 8-byte mbarrier word
void mbar_init(uint64_t *bar, uint32_t count);
void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes);
void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar);
void mbar_wait(uint64_t *bar, uint32_t parity);
}
