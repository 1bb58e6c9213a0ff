// bulk_copy.cuh -- mbarrier + cp.async.bulk (TMA bulk copy, global -> shared, no tensor map) helpers shared by the kernels that
// stage contiguous AoSoA-32 records in shared memory (wilson_dslash3.cu: spinor window; mrhs.cu: link records).
// One elected thread arms the barrier with the expected byte count and issues the copies; consumers wait on the phase parity.
// Sizes and addresses must be multiples of 16 bytes (the records are: 512 B per component row).
#pragma once
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#ifdef LQCD_MBAR_ASM_LOOP
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "K3_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra K3_DONE;\n"
        "bra K3_WAIT;\n"
        "K3_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#else
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
#endif
}
