// lerc_device.cuh -- device helpers shared by the lerc_b200 kernels (sm_100a).
#pragma once
#include "lerc_internal.h"
#include <cstdint>
#include <cfloat>
#include <climits>
#include <cstring>

namespace lerc {

#define LERC_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
  do { LaunchScope scope_((ctx), #kernel);                                                  \
       kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); (ctx)->kernelLaunches++; } while (0)

constexpr unsigned FULL = 0xffffffffu;

// ---- pixel type traits ------------------------------------------------------------------------
template <class T> struct PixelTraits;
#define LERC_TRAITS(T_, CODE_, FLT_, KEY_)                                    \
  template <> struct PixelTraits<T_> {                                        \
    static constexpr int code = CODE_;                                        \
    static constexpr bool isFloat = FLT_;                                     \
    using Key = KEY_;                                                         \
  };
LERC_TRAITS(int8_t, DT_Char, false, uint32_t)
LERC_TRAITS(uint8_t, DT_Byte, false, uint32_t)
LERC_TRAITS(int16_t, DT_Short, false, uint32_t)
LERC_TRAITS(uint16_t, DT_UShort, false, uint32_t)
LERC_TRAITS(int32_t, DT_Int, false, uint32_t)
LERC_TRAITS(uint32_t, DT_UInt, false, uint32_t)
LERC_TRAITS(float, DT_Float, true, uint32_t)
LERC_TRAITS(double, DT_Double, true, unsigned long long)
#undef LERC_TRAITS

// Order-preserving map value -> unsigned key (so atomicMin/atomicMax implement min/max of T).
__device__ __forceinline__ uint32_t toKey(int8_t v)   { return (uint32_t)((int32_t)v) ^ 0x80000000u; }
__device__ __forceinline__ uint32_t toKey(int16_t v)  { return (uint32_t)((int32_t)v) ^ 0x80000000u; }
__device__ __forceinline__ uint32_t toKey(int32_t v)  { return (uint32_t)v ^ 0x80000000u; }
__device__ __forceinline__ uint32_t toKey(uint8_t v)  { return v; }
__device__ __forceinline__ uint32_t toKey(uint16_t v) { return v; }
__device__ __forceinline__ uint32_t toKey(uint32_t v) { return v; }
__device__ __forceinline__ uint32_t toKey(float v) {
  uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ unsigned long long toKey(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
template <class T> __device__ __forceinline__ T fromKey(typename PixelTraits<T>::Key k);
template <> __device__ __forceinline__ int8_t   fromKey<int8_t>(uint32_t k)   { return (int8_t)(int32_t)(k ^ 0x80000000u); }
template <> __device__ __forceinline__ int16_t  fromKey<int16_t>(uint32_t k)  { return (int16_t)(int32_t)(k ^ 0x80000000u); }
template <> __device__ __forceinline__ int32_t  fromKey<int32_t>(uint32_t k)  { return (int32_t)(k ^ 0x80000000u); }
template <> __device__ __forceinline__ uint8_t  fromKey<uint8_t>(uint32_t k)  { return (uint8_t)k; }
template <> __device__ __forceinline__ uint16_t fromKey<uint16_t>(uint32_t k) { return (uint16_t)k; }
template <> __device__ __forceinline__ uint32_t fromKey<uint32_t>(uint32_t k) { return k; }
template <> __device__ __forceinline__ float    fromKey<float>(uint32_t k)    { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
template <> __device__ __forceinline__ double   fromKey<double>(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

// host-side inverse of toKey (same formulas)
template <class T> inline T fromKeyHost(typename PixelTraits<T>::Key k);
template <> inline int8_t   fromKeyHost<int8_t>(uint32_t k)   { return (int8_t)(int32_t)(k ^ 0x80000000u); }
template <> inline int16_t  fromKeyHost<int16_t>(uint32_t k)  { return (int16_t)(int32_t)(k ^ 0x80000000u); }
template <> inline int32_t  fromKeyHost<int32_t>(uint32_t k)  { return (int32_t)(k ^ 0x80000000u); }
template <> inline uint8_t  fromKeyHost<uint8_t>(uint32_t k)  { return (uint8_t)k; }
template <> inline uint16_t fromKeyHost<uint16_t>(uint32_t k) { return (uint16_t)k; }
template <> inline uint32_t fromKeyHost<uint32_t>(uint32_t k) { return k; }
template <> inline float    fromKeyHost<float>(uint32_t k)    { uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k; float f; memcpy(&f, &b, 4); return f; }
template <> inline double   fromKeyHost<double>(unsigned long long k) { unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k; double d; memcpy(&d, &b, 8); return d; }

template <class T> __device__ __forceinline__ bool isNaNVal(T) { return false; }
template <> __device__ __forceinline__ bool isNaNVal<float>(float v) { return v != v; }
template <> __device__ __forceinline__ bool isNaNVal<double>(double v) { return v != v; }

// ---- bit mask (BitMask.h:48-67): pixel k is bit (7 - k%8) of byte k/8 ---------------------------
__device__ __forceinline__ bool maskBit(const uint8_t* bits, long long k) { return (bits[k >> 3] >> (7 - (k & 7))) & 1; }

// ---- small integer helpers ----------------------------------------------------------------------
__device__ __forceinline__ int bitLength(uint32_t v) { return 32 - __clz(v); }          // 0 for v == 0
__device__ __forceinline__ int countFieldBytes(uint32_t n) { return n < 256 ? 1 : (n < 65536 ? 2 : 4); }
__device__ __forceinline__ uint32_t packedBytes(uint32_t n, int nb) { return (uint32_t)(((unsigned long long)n * nb + 7) >> 3); }

// ---- Lerc2 v2 bit stuffing (BitStuffer2.cpp:292-425): values MSB-first inside little-endian uint32 words, the unused low bytes
// of the last word dropped by shifting that word down.  Stored byte holding stream bit s (it is bit 7 - s % 8 of that byte), or -1.
__device__ __forceinline__ int v2ByteOfStreamBit(uint32_t s, uint32_t n, int nb) {
  const uint32_t total = n * (uint32_t)nb, nWords = (total + 31) >> 5, w = s >> 5;
  const int bitsTail = (int)(total & 31), bytesTail = (bitsTail + 7) >> 3, drop = bytesTail > 0 ? 4 - bytesTail : 0;
  const int jj = 3 - (int)((s & 31) >> 3);
  const int k = (int)(4 * w) + jj - (w == nWords - 1 ? drop : 0);
  return k >= (int)(4 * w) ? k : -1;
}

// ---- fp64 arithmetic exactly as the reference's x86-64 build evaluates it: no contraction --------
// (the translation units are additionally compiled with -fmad=false)
__device__ __forceinline__ double blockMaxVal(double zMin, double zMax, double maxZErr) {     // Lerc2.h:337-341
  double fac = __ddiv_rn(1.0, __dmul_rn(2.0, maxZErr));
  return __dmul_rn(__dsub_rn(zMax, zMin), fac);
}
// (unsigned int)(v + 0.5) as the reference's x86-64 build evaluates it (cvttsd2si to 64 bits, low word): NaN -> 0
__device__ __forceinline__ uint32_t roundToUInt(double v) {
  const double t = __dadd_rn(v, 0.5);
  return t == t ? (uint32_t)t : 0u;
}
__device__ __forceinline__ uint32_t quantizeOne(double x, double zMin, double scale) {       // Lerc2.h:369-373
  return (uint32_t)__dadd_rn(__dmul_rn(__dsub_rn(x, zMin), scale), 0.5);
}

// Smallest type holding the block offset exactly (Lerc2.h:457-542).  Returns the 2-bit code; dtUsed out.
__device__ __forceinline__ bool fitsInt(double z, double lo, double hi) { return z >= lo && z <= hi && z == floor(z); }
__device__ __forceinline__ int reduceOffsetType(double z, int dt, int& dtUsed) {
  int tc = 0;
  switch (dt) {
    case DT_Short:  tc = fitsInt(z, -128, 127) ? 2 : (fitsInt(z, 0, 255) ? 1 : 0); dtUsed = dt - tc; break;
    case DT_UShort: tc = fitsInt(z, 0, 255) ? 1 : 0; dtUsed = dt - 2 * tc; break;
    case DT_Int:    tc = fitsInt(z, 0, 255) ? 3 : (fitsInt(z, -32768, 32767) ? 2 : (fitsInt(z, 0, 65535) ? 1 : 0)); dtUsed = dt - tc; break;
    case DT_UInt:   tc = fitsInt(z, 0, 255) ? 2 : (fitsInt(z, 0, 65535) ? 1 : 0); dtUsed = dt - 2 * tc; break;
    case DT_Float:  tc = fitsInt(z, 0, 255) ? 2 : (fitsInt(z, -32768, 32767) ? 1 : 0); dtUsed = tc == 0 ? dt : (tc == 1 ? DT_Short : DT_Byte); break;
    case DT_Double:
      tc = fitsInt(z, -32768, 32767) ? 3 : (fitsInt(z, -2147483648.0, 2147483647.0) ? 2
           : ((z >= -(double)FLT_MAX && z <= (double)FLT_MAX && (double)(float)z == z) ? 1 : 0));
      dtUsed = tc == 0 ? dt : dt - 2 * tc + 1; break;
    default: dtUsed = dt; break;
  }
  return tc;
}
__device__ __forceinline__ int offsetTypeFromCode(int dt, int tc) {                          // Lerc2.h:528-542
  int r;
  switch (dt) {
    case DT_Short: case DT_Int: r = dt - tc; break;
    case DT_UShort: case DT_UInt: r = dt - 2 * tc; break;
    case DT_Float: r = tc == 0 ? dt : (tc == 1 ? DT_Short : DT_Byte); break;
    case DT_Double: r = tc == 0 ? dt : dt - 2 * tc + 1; break;
    default: r = dt; break;
  }
  return (r >= DT_Char && r <= DT_Double) ? r : DT_Undefined;
}
__device__ __forceinline__ int dtSize(int dt) { return dt <= DT_Byte ? 1 : (dt <= DT_UShort ? 2 : (dt <= DT_Float ? 4 : 8)); }

// Block offset as its stored bytes (little endian), up to 8 of them (Lerc2.h:546-613).
__device__ __forceinline__ unsigned long long offsetBits(double z, int dtUsed) {
  switch (dtUsed) {
    case DT_Char:   return (unsigned long long)(uint8_t)(int8_t)z;
    case DT_Byte:   return (unsigned long long)(uint8_t)z;
    case DT_Short:  return (unsigned long long)(uint16_t)(int16_t)z;
    case DT_UShort: return (unsigned long long)(uint16_t)z;
    case DT_Int:    return (unsigned long long)(uint32_t)(int32_t)z;
    case DT_UInt:   return (unsigned long long)(uint32_t)z;
    case DT_Float:  return (unsigned long long)__float_as_uint((float)z);
    default:        return (unsigned long long)__double_as_longlong(z);
  }
}
__device__ __forceinline__ double offsetFromBits(unsigned long long b, int dtUsed) {         // Lerc2.h:617-681
  switch (dtUsed) {
    case DT_Char:   return (double)(int8_t)(uint8_t)b;
    case DT_Byte:   return (double)(uint8_t)b;
    case DT_Short:  return (double)(int16_t)(uint16_t)b;
    case DT_UShort: return (double)(uint16_t)b;
    case DT_Int:    return (double)(int32_t)(uint32_t)b;
    case DT_UInt:   return (double)(uint32_t)b;
    case DT_Float:  return (double)__uint_as_float((uint32_t)b);
    default:        return __longlong_as_double((long long)b);
  }
}

// unaligned little-endian reads from a byte stream
__device__ __forceinline__ unsigned long long loadBytesLE(const uint8_t* p, int n) {
  unsigned long long v = 0;
  for (int i = 0; i < n; i++) v |= (unsigned long long)p[i] << (8 * i);
  return v;
}

// ---- warp reductions ----------------------------------------------------------------------------
template <class V> __device__ __forceinline__ V shflXor(V v, int m) { return __shfl_xor_sync(FULL, v, m); }
template <> __device__ __forceinline__ int8_t shflXor<int8_t>(int8_t v, int m) { return (int8_t)__shfl_xor_sync(FULL, (int)v, m); }
template <> __device__ __forceinline__ uint8_t shflXor<uint8_t>(uint8_t v, int m) { return (uint8_t)__shfl_xor_sync(FULL, (int)v, m); }
template <> __device__ __forceinline__ int16_t shflXor<int16_t>(int16_t v, int m) { return (int16_t)__shfl_xor_sync(FULL, (int)v, m); }
template <> __device__ __forceinline__ uint16_t shflXor<uint16_t>(uint16_t v, int m) { return (uint16_t)__shfl_xor_sync(FULL, (int)v, m); }

template <class V> __device__ __forceinline__ V warpMin(V v) {
#pragma unroll
  for (int m = 16; m; m >>= 1) { V o = shflXor(v, m); v = o < v ? o : v; }
  return v;
}
template <class V> __device__ __forceinline__ V warpMax(V v) {
#pragma unroll
  for (int m = 16; m; m >>= 1) { V o = shflXor(v, m); v = o > v ? o : v; }
  return v;
}
__device__ __forceinline__ int warpSum(int v) {
#pragma unroll
  for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}

// Fletcher-32 from the partial sums A = SUM c, D = SUM (wordIndex mod 65535) * c over the checksum region of
// length len (see k_fletcher_partial in lerc_mask.cu for the derivation).
__host__ __device__ inline uint32_t fletcherFinish(unsigned long long A, unsigned long long D, long long len) {
  const unsigned long long M = 65535ull, m = (unsigned long long)((len + 1) >> 1);
  A %= M; D %= M;
  unsigned long long s1 = (0xffffull + A) % M;
  unsigned long long s2 = ((0xffffull % M) * ((m + 1) % M) + (m % M) * A + (M - D)) % M;
  if (s1 == 0) s1 = M;
  if (s2 == 0) s2 = M;
  return (uint32_t)((s2 << 16) | s1);
}
__host__ __device__ inline void fletcherHostPartial(const uint8_t* bytes, long long r0, long long n, unsigned long long& A, unsigned long long& D) {
  for (long long i = 0; i < n; i++) {
    const long long r = r0 + i;
    const unsigned long long c = (unsigned long long)bytes[i] << ((r & 1) ? 0 : 8);
    A += c; D = (D + ((unsigned long long)((r >> 1) % 65535) * c) % 65535ull) % 65535ull;
  }
}

}  // namespace lerc
