// membw.cu -- standalone probe of B200 memory-system ceilings used to plan the Dslash kernel:
// read-only / copy bandwidth from HBM (working set >> L2) and from L2 (working set << L2).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_read(const double2 *__restrict__ a, size_t n, double *sink) {
    double acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double2 v0 = __ldg(a + i), v1 = __ldg(a + i + stride), v2 = __ldg(a + i + 2 * stride), v3 = __ldg(a + i + 3 * stride);
        acc += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
    }
    for (; i < n; i += stride) { double2 v = __ldg(a + i); acc += v.x + v.y; }
    if (acc == 1.2345e-300) *sink = acc;
}
__global__ void k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = a[i];
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs=%d L2=%d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
    double *sink; cudaMalloc(&sink, 8);
    size_t sizes_mb[] = {16, 32, 64, 96, 256, 1024, 2048};
    double2 *a, *b; cudaMalloc(&a, (size_t)2048 << 20); cudaMalloc(&b, (size_t)2048 << 20);
    cudaMemset(a, 1, (size_t)2048 << 20); cudaMemset(b, 1, (size_t)2048 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bs : {256, 512}) for (int mult : {4, 8}) {
        for (size_t mb : sizes_mb) {
            size_t n = (mb << 20) / 16;
            int grid = p.multiProcessorCount * mult;
            for (int w = 0; w < 3; w++) k_read<<<grid, bs>>>(a, n, sink);
            int reps = mb <= 96 ? 50 : 10;
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) k_read<<<grid, bs>>>(a, n, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double rd = (double)(mb << 20) * reps / (ms * 1e-3) / 1e9;
            for (int w = 0; w < 3; w++) k_copy<<<grid, bs>>>(a, b, n);
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) k_copy<<<grid, bs>>>(a, b, n);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            double cp = 2.0 * (double)(mb << 20) * reps / (ms * 1e-3) / 1e9;
            printf("bs=%d grid=%dxSM ws=%4zu MB  read %8.1f GB/s   copy(r+w) %8.1f GB/s\n", bs, mult, mb, rd, cp);
        }
    }
    return 0;
}
