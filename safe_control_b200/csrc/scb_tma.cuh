// scb_tma.cuh -- bulk-async (TMA engine) global -> shared copies + mbarrier completion, sm_90+/sm_100a PTX.
//
// The per-agent obstacle block OBS[a] is contiguous (56 M bytes): instead of 7 scalar LDG.64 per row through L1
// (17 sectors per request, cbfqp_kernel<.,4,5> at 1 M agents), one elected lane asks the TMA engine to stream the
// block into shared memory (`cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`, SASS UBLKCP) and the
// lanes read their rows conflict-free from there; completion is signalled on an mbarrier the consumers wait on
// (SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK).  A 1-D bulk copy needs no tensor map: 16-byte aligned addresses, size a
// multiple of 16 bytes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace scb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// make the freshly initialised barrier visible to the async proxy (the TMA engine)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy accesses to shared memory (the lanes' reads of the previous tile) are ordered before async-proxy writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// block until the barrier's phase with the given parity has completed (try_wait suspends the thread in hardware)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

}  // namespace scb
