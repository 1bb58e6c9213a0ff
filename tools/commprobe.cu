// commprobe.cu -- standalone probe (single process, 2 GPUs) for the latencies that bound the multi-GPU Dslash:
//   1. ping-pong of a sequence flag over NVLink with st.release.sys / ld.acquire.sys (one-way flag latency)
//   2. per-call cost of ld.acquire.sys vs ld.relaxed.sys on a LOCAL flag that is already set (the face-CTA poll)
//   3. cost of __threadfence_system() / __threadfence() after 6 KB of peer stores per CTA (the pack publish)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/commprobe tools/commprobe.cu ; run on a >= 2 GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void st_rel(unsigned long long *p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long *p) { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_rlx(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

// rank `me` of a 2-party ping-pong: wait for my flag == 2*i+me... then write the peer's flag
__global__ void pingpong(unsigned long long *mine, unsigned long long *peer, int me, int iters, long long *cycles) {
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (me == 0) { st_rel(peer, 2ull * i + 1); while (ld_acq(mine) < 2ull * i + 2) {} }
        else         { while (ld_acq(mine) < 2ull * i + 1) {} st_rel(peer, 2ull * i + 2); }
    }
    *cycles = clock64() - t0;
}
__global__ void poll_cost(const unsigned long long *flag, int n, int relaxed, long long *cycles, unsigned long long *sink) {
    unsigned long long acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) acc += relaxed ? ld_rlx(flag) : ld_acq(flag);
    *cycles = clock64() - t0;
    *sink = acc;
}
__global__ void fence_cost(double2 *dst, const double2 *src, int sys, long long *cycles) {
    // every CTA: 384 complex (6 KB) of stores, bar.sync, thread 0 fences; reports the mean cycles of the fence
    const size_t base = (size_t)blockIdx.x * 384;
    for (int i = threadIdx.x; i < 384; i += blockDim.x) dst[base + i] = src[base + i];
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        if (sys) __threadfence_system(); else __threadfence();
        atomicAdd((unsigned long long *)cycles, (unsigned long long)(clock64() - t0));
    }
}
int main() {
    int nd = 0; cudaGetDeviceCount(&nd);
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    unsigned long long *f[2]; long long *cyc[2]; double2 *buf[2];
    for (int d = 0; d < 2; d++) {
        cudaSetDevice(d); cudaDeviceEnablePeerAccess(1 - d, 0);
        cudaMalloc(&f[d], 64); cudaMemset(f[d], 0, 64); cudaMalloc(&cyc[d], 64); cudaMalloc(&buf[d], 64 << 20);
        cudaMemset(buf[d], 1, 64 << 20);
    }
    const int iters = 2000;
    cudaSetDevice(1); pingpong<<<1, 1>>>(f[1], f[0], 1, iters, cyc[1]);
    cudaSetDevice(0); pingpong<<<1, 1>>>(f[0], f[1], 0, iters, cyc[0]);
    cudaDeviceSynchronize(); cudaSetDevice(1); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc[0], 8, cudaMemcpyDeviceToHost);
    printf("flag ping-pong: %.0f cycles per round trip (%.2f us at 1.9 GHz) -> one-way ~%.2f us\n", (double)c / iters, c / iters / 1900.0, c / iters / 3800.0);
    cudaSetDevice(0);
    unsigned long long *sink; cudaMalloc(&sink, 8);
    for (int rl = 0; rl < 2; rl++) {
        poll_cost<<<1, 1>>>(f[0], 1000, rl, cyc[0], sink); cudaDeviceSynchronize();
        cudaMemcpy(&c, cyc[0], 8, cudaMemcpyDeviceToHost);
        printf("%s poll of a local flag: %.0f cycles per load\n", rl ? "ld.relaxed.sys" : "ld.acquire.sys", c / 1000.0);
    }
    for (int peer = 0; peer < 2; peer++) for (int sys = 0; sys < 2; sys++) {
        cudaMemset(cyc[0], 0, 8);
        const int ctas = 512;
        fence_cost<<<ctas, 128>>>(peer ? buf[1] : buf[0] + (8 << 20) / 16, buf[0], sys, cyc[0]); cudaDeviceSynchronize();
        cudaMemcpy(&c, cyc[0], 8, cudaMemcpyDeviceToHost);
        printf("%s fence after 6 KB of %s stores: %.0f cycles per CTA (512 CTAs)\n", sys ? "__threadfence_system" : "__threadfence", peer ? "PEER" : "local", (double)c / ctas);
    }
    return 0;
}
