// lerc_tma.cuh -- bulk asynchronous copies (TMA, SASS: UBLKCP) and mbarrier helpers for sm_100a.
//
// The raster tiles of the encoder and the stream regions of the decoder are staged in shared memory by the
// copy engine: one elected thread issues `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`
// per row segment and arms an mbarrier with the byte total; the CTA's threads wait on the barrier's phase.
// Requirements of the instruction: 16-byte aligned source, destination and size.
//
// tools/cusim (LERC_CUSIM) has no asynchronous proxy: there the copy is a memcpy done by the issuing thread
// and the wait is the CTA barrier that follows the issue in every kernel that uses these helpers.
#pragma once
#include <cstdint>
#include <cstring>

namespace lerc {

#ifndef LERC_CUSIM

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");     // visible to the async proxy
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smemAddr(smemDst)), "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LERC_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LERC_MBAR_DONE;\n"
      "bra LERC_MBAR_WAIT;\n"
      "LERC_MBAR_DONE:\n"
      "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}

#else   // ---- simulator: synchronous stand-ins ------------------------------------------------------

__device__ __forceinline__ void mbarInit(uint64_t*, uint32_t) {}
__device__ __forceinline__ void mbarExpectTx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t*) { std::memcpy(smemDst, gmemSrc, bytes); }
__device__ __forceinline__ void mbarWait(uint64_t*, uint32_t) {}

#endif

}  // namespace lerc
