// TMA (cp.async.bulk.tensor) + mbarrier helpers for staging u8 image tiles into shared memory.
//
// Every pyramid level of every slot is one 3-D tensor {x: w bytes, y: h rows (stride = 128-B pitch),
// z: slots (stride = slot_bytes)}; a tile load is ONE instruction issued by one thread, the box
// lands row-major in shared memory (no swizzle), out-of-image parts are zero-filled by the
// hardware, and all threads wait on an mbarrier that the copy completes with its byte count.
#pragma once
#include <cuda.h>
#include <cstdint>

namespace b200vo {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barrier visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TMA_DONE;\n"
        "bra TMA_WAIT;\n"
        "TMA_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// dst must be 128-byte aligned; coordinates are in elements (bytes here) and may be negative / past
// the end, but the INNERMOST coordinate must be a multiple of 16 bytes (measured on B200: x = 16 loads,
// x = 27 raises "illegal instruction"; tools/tma_probe) -- callers align the box and keep a shift
__device__ __forceinline__ void load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

}  // namespace tma
}  // namespace b200vo
