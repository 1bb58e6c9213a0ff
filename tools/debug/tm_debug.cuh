// tm_debug.cuh -- DEVELOPMENT AID (not part of the product build): bounded mbarrier waits for wilson_tmarch.cu.
// Build with LQCD_BUILD_DEFS="-DTM_DEBUG" LQCD_BUILD_SUFFIX=_dbg; a wait that does not complete within ~1 s records where it
// happened and the raw barrier words, sets a CTA-wide abort flag (all later waits fall through) and the launcher prints the record.
#pragma once
__device__ unsigned long long tm_dbg[32];
__device__ __forceinline__ unsigned tm_try(uint64_t *bar, uint32_t parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ unsigned long long tm_raw(uint64_t *bar) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(bar)) : "memory");
    return v;
}
__device__ __forceinline__ void tm_wait_dbg(uint64_t *bars, int idx, uint32_t parity, int code, int r, int task, volatile int *abort_flag) {
    if (*abort_flag) return;
    const long long t0 = clock64();
    while (!tm_try(&bars[idx], parity)) {
        if (*abort_flag) return;
        if (clock64() - t0 > 2000000000LL) {
            *abort_flag = 1;
            if (atomicAdd(&tm_dbg[0], 1ull) == 0ull) {
                tm_dbg[1] = (unsigned long long)code; tm_dbg[2] = blockIdx.x; tm_dbg[3] = threadIdx.x; tm_dbg[4] = (unsigned long long)r;
                tm_dbg[5] = (unsigned long long)task; tm_dbg[6] = parity; tm_dbg[7] = (unsigned long long)idx;
                for (int j = 0; j < 7; j++) tm_dbg[8 + j] = tm_raw(&bars[j]);
            }
            return;
        }
    }
}
