// lerc_fletcher.cuh -- Fletcher-32 partial sums of 16-byte chunks with dp4a (Lerc2::ComputeChecksumFletcher32, Lerc2.cpp:1037-1064;
// closed form: SURVEY.md Appendix B.9, finished by fletcherFinish in lerc_device.cuh).  Shared by the encoder's flush and the
// decoder's stream pass.
#pragma once
#include "lerc_device.cuh"

namespace lerc {

// Fletcher-32 partial sums of one 16-byte output chunk whose first byte has checksum-region offset r0 (parity PAR):
// S = sum of the chunk's bytes weighted 256 (even region offsets) / 1 (odd), S1 = the same weighted with the byte's
// word index relative to the chunk's first word.
template <int PAR>
__device__ __forceinline__ void fletcherChunk(const uint32_t (&o)[4], uint32_t& S, uint32_t& S1) {
  uint32_t H = 0, L = 0, HW = 0, LW = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (PAR == 0) {   // bytes 0, 2 of a word are high bytes; word index of byte 4j + m: (4j + m) >> 1
      H = __dp4a(o[j], 0x00010001u, H);  L = __dp4a(o[j], 0x01000100u, L);
      HW = __dp4a(o[j], (uint32_t)(2 * j) | ((uint32_t)(2 * j + 1) << 16), HW);
      LW = __dp4a(o[j], ((uint32_t)(2 * j) << 8) | ((uint32_t)(2 * j + 1) << 24), LW);
    } else {          // bytes 1, 3 are high bytes; word index of byte 4j + m: (4j + m + 1) >> 1
      H = __dp4a(o[j], 0x01000100u, H);  L = __dp4a(o[j], 0x00010001u, L);
      HW = __dp4a(o[j], ((uint32_t)(2 * j + 1) << 8) | ((uint32_t)(2 * j + 2) << 24), HW);
      LW = __dp4a(o[j], (uint32_t)(2 * j) | ((uint32_t)(2 * j + 1) << 16), LW);
    }
  }
  S = 256u * H + L; S1 = 256u * HW + LW;
}

// x mod 65535 without a 64-bit division: 2^16 == 1 (mod 65535), so the 16-bit digits of x may simply be added
__device__ __forceinline__ uint32_t mod65535(unsigned long long x) {
  uint32_t s = (uint32_t)(x & 0xffff) + (uint32_t)((x >> 16) & 0xffff) + (uint32_t)((x >> 32) & 0xffff) + (uint32_t)(x >> 48);   // < 2^18
  s = (s & 0xffff) + (s >> 16);                                            // < 2^16 + 4
  s = (s & 0xffff) + (s >> 16);
  return s == 65535u ? 0u : s;
}
// fletcherFinish (lerc_device.cuh) for device code that runs on a single warp at the end of a kernel: same value, 32-bit arithmetic
__device__ __forceinline__ uint32_t fletcherFinishFast(unsigned long long A, unsigned long long D, long long len) {
  const uint32_t M = 65535u;
  const uint32_t a = mod65535(A), d = mod65535(D), m = mod65535((unsigned long long)((len + 1) >> 1));
  uint32_t s1 = mod65535((unsigned long long)a);                            // (0xffff + a) mod M == a mod M
  uint32_t s2 = mod65535((unsigned long long)m * a + (M - d));             // (0xffff mod M) * (m + 1) == 0
  if (s1 == 0) s1 = M;
  if (s2 == 0) s2 = M;
  return (s2 << 16) | s1;
}

}  // namespace lerc
