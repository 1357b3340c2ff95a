// lerc_tma.cuh -- bulk asynchronous copies (TMA, SASS: UBLKCP) and mbarrier helpers for sm_100a.
//
// The raster tiles of the encoder and the stream regions of the decoder are staged in shared memory by the
// copy engine: one elected thread issues `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`
// per row segment and arms an mbarrier with the byte total; the CTA's threads wait on the barrier's phase.
// Requirements of the instruction: 16-byte aligned source, destination and size.
//
// tools/cusim (LERC_CUSIM) has no asynchronous proxy: there the copy is a memcpy done by the issuing thread, the
// barrier word is a plain phase counter and a wait yields to the CTA's other threads until the phase has flipped.
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
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
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
// barrier among the first `nThreads` threads' worth of warps that name `id` (1..15; 0 is __syncthreads)
__device__ __forceinline__ void namedBarSync(int id, int nThreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nThreads) : "memory"); }

#else   // ---- simulator: the barrier word counts completed phases (low 32 bits) and pending arrivals / bytes -----------

// word layout in the simulator: bits 0..15 completed phases, 16..31 arrivals still expected in this phase (reloaded from
// bits 32..47), 48..63 unused; the byte count of bulk copies is not modelled: a copy is a memcpy by the issuing thread.
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t arrivals) { *bar = ((uint64_t)arrivals << 32) | ((uint64_t)arrivals << 16); }
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
  uint64_t v = *bar;
  uint32_t pending = (uint32_t)(v >> 16) & 0xffff, phases = (uint32_t)v & 0xffff, init = (uint32_t)(v >> 32) & 0xffff;
  if (--pending == 0) { phases = (phases + 1) & 0xffff; pending = init; }
  *bar = ((uint64_t)init << 32) | ((uint64_t)pending << 16) | phases;
}
// expect_tx + arrive: in the simulator the copies that follow are synchronous, so the phase may only complete after them:
// the arrival is recorded by mbarSimCopiesDone(), which every caller of mbarExpectTx invokes after its last bulkLoad.
__device__ __forceinline__ void mbarExpectTx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t*) { std::memcpy(smemDst, gmemSrc, bytes); }
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  while ((((uint32_t)*(volatile uint64_t*)bar) & 1u) == parity) __nanosleep(0);   // yields to the CTA's other fibers
}
__device__ __forceinline__ void namedBarSync(int id, int nThreads) { cusim::namedBarrier(id, nThreads); }

#endif

// after the last bulkLoad of an expect_tx group (GPU: nothing, the copy engine completes the phase; simulator: the arrival)
__device__ __forceinline__ void mbarSimCopiesDone(uint64_t* bar) {
#ifdef LERC_CUSIM
  mbarArrive(bar);
#else
  (void)bar;
#endif
}

}  // namespace lerc
