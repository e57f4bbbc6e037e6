// rdr_device.cuh -- device-only helpers: TMA bulk staging of the scene blob into shared memory.
#pragma once

#include "rdr_trace.cuh"

namespace rdr {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// global blob -> shared memory with one cp.async.bulk (SASS UBLKCP) completing on an mbarrier;
// thread 0 issues, every thread waits on phase 0.  bytes is a multiple of 16, both sides 16-byte aligned.
__device__ __forceinline__ void stage_blob(unsigned char *dst, const unsigned char *src, uint32_t bytes, uint64_t *bar)
{
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a), "r"(0u) : "memory");
    }
}

// dynamic shared memory: [blob][mbarrier (16 B)][masks: max_chunks * blockDim words]
__device__ __forceinline__ uint64_t *bar_ptr(unsigned char *smem, const SceneLayout &L)
{
    return reinterpret_cast<uint64_t *>(smem + L.blob_bytes);
}
__device__ __forceinline__ uint32_t *mask_base(unsigned char *smem, const SceneLayout &L)
{
    return reinterpret_cast<uint32_t *>(smem + L.blob_bytes + 16u);
}

}  // namespace rdr
