/*
 * lerc_b200.h -- entry points lerc_b200 adds next to the reference C API (Lerc_c_api.h).
 * Plain C ABI: pointers and sizes only, no CUDA or torch types in the signatures.
 *
 * Nothing here replaces a reference interface; these are the additive extensions announced in
 * (i) the 12 reference functions accept CUDA device pointers for every
 * buffer argument (no new symbol needed), (ii) the caller may supply the CUDA stream the kernels run
 * on, so device-resident pipelines can be timed and ordered with CUDA events, (iii) counters that
 * let a harness prove the CUDA path ran (number of kernel launches, calls).
 */
#ifndef LERC_B200_H
#define LERC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
  #define LERC_B200_API __attribute__((visibility("default")))
#else
  #define LERC_B200_API
#endif

/* Device pointers and ordering: without lerc_b200_set_stream the library works on a private NON-BLOCKING stream, so
 * device buffers handed to lerc_* must be complete (synchronise the producing stream first); results are complete
 * when the call returns.  With lerc_b200_set_stream the calls are ordered on the given stream like any other work.
 *
 * Use `cudaStream` (a cudaStream_t passed as void*) for all work issued by the calling thread's next
 * lerc_* calls when enable != 0; enable == 0 returns to the library's private non-blocking stream.
 * Thread-local.  The lerc_* calls still return only after their results are complete. */
LERC_B200_API void lerc_b200_set_stream(void* cudaStream, int enable);

/* out[0] = kernels launched so far, out[1] = encode calls, out[2] = decode calls,
 * out[3] = encodes that took the fused all-valid fast path, out[4] = decodes that did. */
LERC_B200_API void lerc_b200_get_stats(unsigned long long* out, int n);

/* Per-kernel timing: while enabled, every kernel launch of this library is bracketed by a CUDA event pair on
 * its stream.  lerc_b200_get_profile() writes lines "name<TAB>launches<TAB>total_ms" into buf. */
LERC_B200_API void lerc_b200_profile(int enable);
LERC_B200_API void lerc_b200_get_profile(char* buf, int bufLen, int reset);

LERC_B200_API const char* lerc_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
