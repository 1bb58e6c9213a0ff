// p2pbw.cu -- probe of NVLink peer-store performance from a kernel (single process, 2 devices):
// bandwidth of 16-byte-per-lane coalesced stores into the peer's memory for halo-sized and large buffers,
// and the cost of the "fence + flag" publication.  Used to size the halo pack kernel.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_store(double2 *__restrict__ dst, const double2 *__restrict__ src, size_t n, int fence) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
    if (fence) { __syncthreads(); if (threadIdx.x == 0) __threadfence_system(); }
}
int main() {
    int nd = 0; cudaGetDeviceCount(&nd);
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    int can = 0; cudaDeviceCanAccessPeer(&can, 0, 1); printf("peer access 0->1: %d\n", can);
    cudaSetDevice(1); double2 *remote; cudaMalloc(&remote, (size_t)256 << 20);
    cudaSetDevice(0); cudaDeviceEnablePeerAccess(1, 0);
    double2 *local, *local2; cudaMalloc(&local, (size_t)256 << 20); cudaMalloc(&local2, (size_t)256 << 20);
    cudaMemset(local, 1, (size_t)256 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t sizes[] = {(size_t)1 << 20, (size_t)3 << 20, (size_t)6 << 20, (size_t)24 << 20, (size_t)256 << 20};
    for (int fence = 0; fence < 2; fence++)
    for (size_t bytes : sizes) {
        for (int grid : {148, 592, 2368}) {
            size_t n = bytes / 16;
            for (int tgt = 0; tgt < 2; tgt++) {
                double2 *dst = tgt ? remote : local2;
                for (int w = 0; w < 3; w++) k_store<<<grid, 256>>>(dst, local, n, fence);
                cudaEventRecord(e0);
                const int reps = 20;
                for (int r = 0; r < reps; r++) k_store<<<grid, 256>>>(dst, local, n, fence);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("fence=%d %6.1f MB grid=%4d %s: %7.2f us/kernel  %7.1f GB/s\n", fence, bytes / 1048576.0, grid,
                       tgt ? "peer " : "local", 1e3 * ms / reps, bytes * reps / (ms * 1e-3) / 1e9);
            }
        }
    }
    return 0;
}
