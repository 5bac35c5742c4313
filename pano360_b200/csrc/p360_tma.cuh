// Bulk asynchronous copies global -> shared memory (the TMA engine, 1-D form) completing on an
// mbarrier: cp.async.bulk + mbarrier.try_wait (SASS: UBLKCP / SYNCS).  Used by the vertical blur,
// whose staged tile is a stack of contiguous 512-byte row segments: one copy per row replaces a
// load + address arithmetic + store per thread and element, and the threads that would have
// staged go straight to waiting on the barrier.  sm_100a only; the host build of the kernels
// (tests/emul) defines P360_EMUL_BUILD and keeps the per-thread staging.
#pragma once
#include <stdint.h>

namespace p360 {

#ifdef P360_EMUL_BUILD
constexpr bool kHaveTma = false;
#else
constexpr bool kHaveTma = true;
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
#ifdef P360_EMUL_BUILD
    return 0u;
#else
    return (uint32_t)__cvta_generic_to_shared(p);
#endif
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// make the initialised barrier (and prior generic-proxy accesses to shared memory) visible to the async proxy
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bytes: a multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_global, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

}  // namespace p360
