// lerc_encode.cu -- Lerc2 v6 band encoder on the GPU.
//
// Pipeline of one band (reference: Lerc::EncodeInternal Lerc.cpp:628-789, Lerc2::ComputeNumBytesNeededToWrite
// Lerc2.cpp:179-381, Lerc2::Encode :396-480):
//   mask build (+NaN fold)  ->  per-depth min/max (+all-integer test)  ->  [float: raise maxZError]
//   ->  micro-block count pass (8x8)  ->  [8-bit lossless: histograms -> Huffman tables on the host]
//   ->  [16x16 retry]  ->  one-sweep decision  ->  header/mask/ranges  ->  block write pass | Huffman | raw
//   ->  Fletcher-32.
// The host only takes the image-global decisions (a handful of scalars per band); every per-pixel and
// per-block step is a kernel below.  Reference citations relative to /root/reference/src/LercLib.
#include "lerc_device.cuh"
#include "lerc_kernels.h"
#include <cub/device/device_scan.cuh>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <atomic>
#include <cstdlib>
#include <chrono>

namespace lerc {

// =================================================================================================
// prefix sums (CUB; plumbing, not a hot kernel)

void exclusiveScanU32(Context* ctx, const uint32_t* dIn, uint32_t* dOut, size_t n) {
  LaunchScope scope(ctx, "cub::ExclusiveSum<u32>");
  size_t tmpBytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, dIn, dOut, (int)(n + 1), ctx->stream);
  void* tmp = ctx->arena.alloc(tmpBytes ? tmpBytes : 16);
  cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, dIn, dOut, (int)(n + 1), ctx->stream);
  ctx->kernelLaunches += 2;
}
void exclusiveScanU64(Context* ctx, const unsigned long long* dIn, unsigned long long* dOut, size_t n) {
  LaunchScope scope(ctx, "cub::ExclusiveSum<u64>");
  size_t tmpBytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, dIn, dOut, (int)(n + 1), ctx->stream);
  void* tmp = ctx->arena.alloc(tmpBytes ? tmpBytes : 16);
  cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, dIn, dOut, (int)(n + 1), ctx->stream);
  ctx->kernelLaunches += 2;
}

// =================================================================================================
// per-depth data ranges (+ float extras)              Lerc2.cpp:1404-1470, Lerc.cpp:1420-1500
//
// "First occurrence wins among equal values" only matters for +-0.0: besides the
// order-preserving min/max keys each depth tracks the first pixel holding a zero; the finish kernel takes
// the sign from there.

enum { STATF_NAN = 1, STATF_NOT_INT = 2 };

template <class T> struct StatsBuffers {
  typename PixelTraits<T>::Key* minKey;   // [nDepth]
  typename PixelTraits<T>::Key* maxKey;   // [nDepth]
  uint32_t* zeroIdx;                      // [nDepth], float types only
  int* flags;
  double* ranges;                         // [2*nDepth]: mins then maxs (finish kernel)
};

template <class K> __device__ __forceinline__ K keyMaxValue();
template <> __device__ __forceinline__ uint32_t keyMaxValue<uint32_t>() { return 0xffffffffu; }
template <> __device__ __forceinline__ unsigned long long keyMaxValue<unsigned long long>() { return ~0ull; }

template <class T>
__global__ void k_stats_init(StatsBuffers<T> sb, int nDepth) {
  using K = typename PixelTraits<T>::Key;
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < nDepth; d += gridDim.x * blockDim.x) {
    sb.minKey[d] = keyMaxValue<K>(); sb.maxKey[d] = 0; sb.zeroIdx[d] = 0xffffffffu;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *sb.flags = 0;
}

// Each block owns a contiguous pixel range; when nDepth <= 256 a thread always sees the same depth, so
// its min/max live in registers and meet the other threads' through shared-memory atomics once.
template <class T>
__global__ void k_stats(const T* __restrict__ data, const uint8_t* __restrict__ bits, long long nPix, int nDepth,
                        int pixPerBlock, StatsBuffers<T> sb) {
  using K = typename PixelTraits<T>::Key;
  __shared__ K sMin[256], sMax[256];
  __shared__ uint32_t sZero[256];
  const long long p0 = (long long)blockIdx.x * pixPerBlock;
  const long long p1 = p0 + pixPerBlock < nPix ? p0 + pixPerBlock : nPix;
  const int depthBase = blockIdx.y * 256;
  const int nd = nDepth - depthBase < 256 ? nDepth - depthBase : 256;     // depths handled by this block
  if ((int)threadIdx.x < nd) { sMin[threadIdx.x] = keyMaxValue<K>(); sMax[threadIdx.x] = 0; sZero[threadIdx.x] = 0xffffffffu; }
  __syncthreads();
  const int used = (256 / nd) * nd;
  int flags = 0;
  if ((int)threadIdx.x < used) {
    const int slot = threadIdx.x % nd, d = depthBase + slot;
    const int pixStep = used / nd;
    K mn = keyMaxValue<K>(), mx = 0;
    uint32_t zi = 0xffffffffu;
    for (long long p = p0 + threadIdx.x / nd; p < p1; p += pixStep) {
      if (bits && !maskBit(bits, p)) continue;
      const T v = data[p * nDepth + d];
      if (PixelTraits<T>::isFloat) {
        if (isNaNVal(v)) { flags |= STATF_NAN; continue; }
        if (!(v == (T)floor((double)v + 0.5))) flags |= STATF_NOT_INT;             // Lerc.h:248
        if (v == (T)0 && (uint32_t)p < zi) zi = (uint32_t)p;
      }
      const K k = toKey(v);
      mn = k < mn ? k : mn; mx = k > mx ? k : mx;
    }
    if (mx >= mn) { atomicMin(&sMin[slot], mn); atomicMax(&sMax[slot], mx); }
    if (zi != 0xffffffffu) atomicMin(&sZero[slot], zi);
  }
  flags = __reduce_or_sync(FULL, flags);
  if (flags && (threadIdx.x & 31) == 0) atomicOr(sb.flags, flags);
  __syncthreads();
  if ((int)threadIdx.x < nd && sMax[threadIdx.x] >= sMin[threadIdx.x]) {
    const int d = depthBase + threadIdx.x;
    atomicMin(&sb.minKey[d], sMin[threadIdx.x]); atomicMax(&sb.maxKey[d], sMax[threadIdx.x]);
    if (sZero[threadIdx.x] != 0xffffffffu) atomicMin(&sb.zeroIdx[d], sZero[threadIdx.x]);
  }
}

template <class T>
__global__ void k_stats_finish(const T* __restrict__ data, int nDepth, StatsBuffers<T> sb) {
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < nDepth; d += gridDim.x * blockDim.x) {
    T lo = fromKey<T>(sb.minKey[d]), hi = fromKey<T>(sb.maxKey[d]);
    if (PixelTraits<T>::isFloat && sb.zeroIdx[d] != 0xffffffffu) {
      const T z = data[(long long)sb.zeroIdx[d] * nDepth + d];      // the first zero in scan order carries the sign
      if (lo == (T)0) lo = z;
      if (hi == (T)0) hi = z;
    }
    sb.ranges[d] = (double)lo; sb.ranges[nDepth + d] = (double)hi;
  }
}

// =================================================================================================
// float data already on a coarser decimal grid: raise maxZError                  Lerc2.cpp:1233-1339
//
// For each candidate factor f the reference tracks max |round(x*f) - x*f| over the valid values, skipping
// (break) the finer candidates of a value that is exactly on a coarser grid.  x*f is exact in fp64 (24/53-bit
// significand times <= 14 bits for float; for double the product is rounded once, as in the reference), and a
// value on a coarser grid is also on every finer one (each factor divides the next), where its rounding
// distance is 0.  So the per-candidate maximum over ALL valid values equals the reference's bookkeeping, and
// row-wise pruning only removes candidates whose final maximum would fail anyway.
// bit-plane mode (Lerc2::TryBitPlaneCompression, Lerc2.cpp:1071-1229): per bit plane, how often do horizontally / vertically
// neighbouring valid pixels differ in that bit?  grid.y = depth; counts[depth * 32 + plane], pairs = number of neighbour pairs.
// allValid1: the reference's special case (nDepth == 1, every pixel valid) leaves out the last row and the last column.
template <class T>
__global__ void k_bitplane_counts(const T* __restrict__ data, const uint8_t* __restrict__ bits, int nRows, int nCols, int nDepth, int allValid1,
                                  unsigned long long* __restrict__ counts, unsigned long long* __restrict__ pairs) {
  constexpr int NB = 8 * (int)sizeof(T);
  const int m = blockIdx.y;
  unsigned int c[NB];
#pragma unroll
  for (int s = 0; s < NB; s++) c[s] = 0;
  unsigned long long np = 0;
  const long long nPix = (long long)nRows * nCols;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nPix; k += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(k / nCols), j = (int)(k - (long long)i * nCols);
    bool hori, vert;
    if (allValid1) { hori = vert = (i < nRows - 1 && j < nCols - 1); }
    else {
      const bool v = !bits || maskBit(bits, k);                       // bits == nullptr: every pixel valid (nDepth > 1)
      hori = v && j < nCols - 1 && (!bits || maskBit(bits, k + 1));
      vert = v && i < nRows - 1 && (!bits || maskBit(bits, k + nCols));
    }
    const long long m0 = k * nDepth + m;
    const uint32_t x = (uint32_t)(long long)data[m0];
    if (hori) {
      const uint32_t d = x ^ (uint32_t)(long long)data[m0 + nDepth];
#pragma unroll
      for (int s = 0; s < NB; s++) c[s] += (d >> s) & 1u;
      np++;
    }
    if (vert) {
      const uint32_t d = x ^ (uint32_t)(long long)data[m0 + (long long)nDepth * nCols];
#pragma unroll
      for (int s = 0; s < NB; s++) c[s] += (d >> s) & 1u;
      np++;
    }
  }
#pragma unroll
  for (int s = 0; s < NB; s++) {
    unsigned int v = c[s];
    for (int sh = 16; sh; sh >>= 1) v += __shfl_xor_sync(FULL, v, sh);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counts[(size_t)m * 32 + s], (unsigned long long)v);
  }
  for (int sh = 16; sh; sh >>= 1) np += __shfl_xor_sync(FULL, np, sh);
  if (m == 0 && (threadIdx.x & 31) == 0 && np) atomicAdd(pairs, np);
}

// codec versions 2..5: NaN -> -FLT_MAX / -DBL_MAX (Lerc::ReplaceNaNValues, Lerc.cpp:906-930); pixels whose depths are all NaN
// were already made invalid by k_mask_build, their values no longer matter
template <class T>
__global__ void k_replace_nan(T* __restrict__ data, long long nElem) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nElem; e += (long long)gridDim.x * blockDim.x)
    if (isNaNVal(data[e])) data[e] = (T)(sizeof(T) == 4 ? -FLT_MAX : -DBL_MAX);
}

struct RaiseArgs { double fac[9]; int n; };

template <class T>
__global__ void k_try_raise(const T* __restrict__ data, const uint8_t* __restrict__ bits, long long p0, long long p1, int nDepth,
                            RaiseArgs ra, unsigned long long* __restrict__ maxBits /*[9], zeroed*/) {
  double m[9];
#pragma unroll
  for (int c = 0; c < 9; c++) m[c] = 0;
  const long long e0 = p0 * nDepth, e1 = p1 * nDepth;
  for (long long e = e0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += (long long)gridDim.x * blockDim.x) {
    if (bits && !maskBit(bits, e / nDepth)) continue;
    const double x = (double)data[e];
#pragma unroll
    for (int c = 0; c < 9; c++) {
      if (c < ra.n) {
        const double z = __dmul_rn(x, ra.fac[c]);
        const double dlt = fabs(__dsub_rn(floor(__dadd_rn(z, 0.5)), z));
        if (dlt > m[c]) m[c] = dlt;           // NaN / Inf never win, as with std::max(a, NaN)
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 9; c++) {
    if (c < ra.n) {
      unsigned long long b = (unsigned long long)__double_as_longlong(m[c]);     // non-negative doubles order like integers
      for (int s = 16; s; s >>= 1) { unsigned long long o = __shfl_xor_sync(FULL, b, s); b = o > b ? o : b; }
      if ((threadIdx.x & 31) == 0 && b) atomicMax(&maxBits[c], b);
    }
  }
}

// =================================================================================================
// micro-block coder: one warp per block position, all depths                  Lerc2.cpp:1474-1668
//
// The warp compacts the block's valid pixels into shared memory (row-major order, like
// GetValidDataAndStats, Lerc2.cpp:1717-1799), reduces min/max/same-value count, sizes the candidate
// codings (NumBytesTile, Lerc2.h:416-453) and, in the write pass, emits the bytes of WriteTile
// (Lerc2.cpp:1949-2021) / BitStuffer2::EncodeSimple / EncodeLut (BitStuffer2.cpp:35-153) at the block's
// scanned offset.  This is the general path: masks, edge blocks, nDepth > 1 with depth-delta blocks,
// LUT blocks, both block sizes, all 8 pixel types.

struct TileArgs {
  const void* data; const uint8_t* bits;          // bits == nullptr: every pixel valid
  int nRows, nCols, nDepth, mb, nTx, nTy, dt, version;
  double maxZErr; uint32_t maxQ;
  int tryDiff, checkOverflow, allValidImage;
  int onlyFlagged;                                // count pass: only the blocks whose blockBytes entry is 0xffffffff (left over by k_tiles_count8)
  uint32_t* blockBytes;                           // count pass: out [nBlocks]
  const uint32_t* blockOff;                       // write pass: in  [nBlocks], exclusive prefix of blockBytes
  uint8_t* out;                                   // write pass: start of the block stream
};

enum { BEM_RAW = 0, BEM_SIMPLE = 1, BEM_LUT = 2, BEM_ZERO = 3 };
struct BlockChoice { int nBytes, mode, nb, nLut, nbIdx, tc, dtUsed; uint32_t maxElem; };

template <class T> struct TileSmem { T* buf; T* prev; int32_t* dbuf; uint32_t* q; uint32_t* qd; uint32_t* rank; uint32_t* lut; };

template <class T> __host__ __device__ inline size_t tileSmemBytesPerWarp(int area) {
  return (size_t)area * (2 * sizeof(T) + 4 * 4) + 256 * 4 + 64;
}
template <class T> __device__ inline TileSmem<T> carveTileSmem(uint8_t* base, int area) {
  TileSmem<T> s;
  uint8_t* p = base;
  s.buf = (T*)p;  p += (size_t)area * sizeof(T);
  s.prev = (T*)p; p += (size_t)area * sizeof(T);
  s.dbuf = (int32_t*)p; p += (size_t)area * 4;
  s.q = (uint32_t*)p;   p += (size_t)area * 4;
  s.qd = (uint32_t*)p;  p += (size_t)area * 4;
  s.rank = (uint32_t*)p; p += (size_t)area * 4;
  s.lut = (uint32_t*)p;
  return s;
}

// n values of `vals`, nb bits each, LSB first, written byte by byte (BitStuffer2.cpp:432-472)
__device__ inline void emitPacked(uint8_t* __restrict__ p, const uint32_t* __restrict__ vals, uint32_t n, int nb, int lane, int version = 6) {
  const uint32_t nBytes = packedBytes(n, nb);
  if (version < 3) {                                   // codec version 2: MSB-first words (compatibility writer, bit by bit)
    const uint32_t total = n * (uint32_t)nb, nWords = (total + 31) >> 5;
    const int bitsTail = (int)(total & 31), bytesTail = (bitsTail + 7) >> 3, drop = bytesTail > 0 ? 4 - bytesTail : 0;
    for (uint32_t k = lane; k < nBytes; k += 32) {
      const uint32_t w = k >> 2;
      const int jj = (int)(k & 3) + (w == nWords - 1 ? drop : 0);          // byte of the little-endian word this stored byte came from
      const uint32_t s0 = 32 * w + (uint32_t)(3 - jj) * 8;                   // stream bit held by bit 7 of the byte
      uint32_t acc = 0;
      for (int u = 0; u < 8; u++) {
        const uint32_t sb = s0 + (uint32_t)u, i = sb / (uint32_t)nb, t = sb - i * (uint32_t)nb;
        if (i < n && ((vals[i] >> (nb - 1 - (int)t)) & 1)) acc |= 0x80u >> u;
      }
      p[k] = (uint8_t)acc;
    }
    return;
  }
  for (uint32_t k = lane; k < nBytes; k += 32) {
    const uint32_t bit0 = k * 8;
    uint32_t i = bit0 / (uint32_t)nb, acc = 0;
    for (; i < n; i++) {
      const int sh = (int)(i * (uint32_t)nb) - (int)bit0;
      if (sh >= 8) break;
      const uint32_t v = vals[i];
      acc |= sh >= 0 ? (v << sh) : (v >> (-sh));
    }
    p[k] = (uint8_t)acc;
  }
}

// distinct-value ranks of q[0..n): rank[i] = number of distinct values smaller than q[i]; returns #distinct.
// lut[r] = the value of rank r.  O(n^2/32) with broadcast shared-memory reads; only LUT candidates pay it.
__device__ inline int rankDistinct(const uint32_t* __restrict__ q, int n, uint32_t* __restrict__ rank, uint32_t* __restrict__ lut, int lane) {
  int nFirst = 0;
  for (int i = lane; i < n; i += 32) {           // pass 1: first occurrence flags (kept in the top bit of rank[])
    const uint32_t qi = q[i];
    bool first = true;
    for (int j = 0; j < i; j++) if (q[j] == qi) { first = false; break; }
    rank[i] = first ? 0x80000000u : 0u;
    nFirst += first ? 1 : 0;
  }
  nFirst = warpSum(nFirst);
  __syncwarp();
  for (int i = lane; i < n; i += 32) {           // pass 2: count distinct smaller values
    const uint32_t qi = q[i];
    uint32_t r = 0;
    for (int j = 0; j < n; j++) r += ((rank[j] & 0x80000000u) && q[j] < qi) ? 1u : 0u;
    rank[i] |= r;
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const uint32_t r = rank[i];
    if ((r & 0x80000000u) && (r & 0xffffu) < 256) lut[r & 0xffffu] = q[i];
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) rank[i] &= 0x7fffffffu;
  __syncwarp();
  return nFirst;
}

// Lerc2.h:416-453 (NumBytesTile) + BitStuffer2::ComputeNumBytesNeeded{Simple,Lut}.  Warp-uniform.
__device__ inline BlockChoice sizeBlock(const TileArgs& a, int n, double zMin, double zMax, int elemSize, int dtZ, bool tryLut,
                                        const uint32_t* q, uint32_t* rank, uint32_t* lut, int lane) {
  BlockChoice c; c.mode = BEM_RAW; c.nb = 0; c.nLut = 0; c.nbIdx = 0; c.tc = 0; c.dtUsed = dtZ; c.maxElem = 0;
  if (n == 0 || (zMin == 0 && zMax == 0)) { c.nBytes = 1; c.mode = BEM_ZERO; return c; }
  const int raw = 1 + n * elemSize;
  double mv = 0;
  if ((a.maxZErr == 0 && zMax > zMin) || (a.maxZErr > 0 && (mv = blockMaxVal(zMin, zMax, a.maxZErr)) > (double)a.maxQ)) { c.nBytes = raw; return c; }
  c.tc = reduceOffsetType(zMin, dtZ, c.dtUsed);
  int nBytes = 1 + dtSize(c.dtUsed);
  c.maxElem = roundToUInt(mv);
  bool useLut = false;
  if (c.maxElem > 0) {
    c.nb = bitLength(c.maxElem);
    const int simple = 1 + countFieldBytes(n) + (int)packedBytes(n, c.nb);
    int best = simple;
    if (tryLut) {
      const int nDistinct = rankDistinct(q, n, rank, lut, lane);
      c.nLut = nDistinct - 1; c.nbIdx = bitLength((uint32_t)c.nLut);
      const int lutBytes = 1 + countFieldBytes(n) + 1 + (int)packedBytes(c.nLut, c.nb) + (int)packedBytes(n, c.nbIdx);
      useLut = lutBytes < simple;
      best = useLut ? lutBytes : simple;
    }
    nBytes += best;
  }
  if (nBytes < raw) c.mode = (useLut && c.maxElem > 0) ? BEM_LUT : BEM_SIMPLE;
  else nBytes = raw;
  c.nBytes = nBytes;
  return c;
}

__device__ inline bool needQuantize(const TileArgs& a, int n, double zMin, double zMax) {     // Lerc2.h:345-353
  if (n == 0 || a.maxZErr == 0) return false;
  const double mv = blockMaxVal(zMin, zMax, a.maxZErr);
  return !(mv > (double)a.maxQ || roundToUInt(mv) == 0);
}

// Lerc2.cpp:1949-2021.  Returns the number of bytes emitted (== choice.nBytes).
template <class V>
__device__ inline int emitBlock(const TileArgs& a, uint8_t* p, const V* vals, int n, int j0, double zMin, const BlockChoice& c,
                                bool diff, const uint32_t* q, const uint32_t* rank, const uint32_t* lut, int lane) {
  uint8_t flag = (uint8_t)(((j0 >> 3) & 15) << 2);
  if (a.version >= 5) flag = diff ? (uint8_t)(flag | 4) : (uint8_t)(flag & 0x38);
  if (c.mode == BEM_ZERO) { if (lane == 0) p[0] = flag | 2; return 1; }
  if (c.mode == BEM_RAW) {
    if (lane == 0) p[0] = flag;
    const uint8_t* src = (const uint8_t*)vals;
    const int nb = n * (int)sizeof(V);
    for (int k = lane; k < nb; k += 32) p[1 + k] = src[k];
    return 1 + nb;
  }
  const int osz = dtSize(c.dtUsed);
  if (lane == 0) {
    p[0] = (uint8_t)(flag | (c.maxElem == 0 ? 3 : 1) | (c.tc << 6));
    const unsigned long long ob = offsetBits(zMin, c.dtUsed);
    for (int k = 0; k < osz; k++) p[1 + k] = (uint8_t)(ob >> (8 * k));
  }
  int pos = 1 + osz;
  if (c.maxElem == 0) return pos;
  const int cb = countFieldBytes(n);
  if (lane == 0) {
    p[pos] = (uint8_t)(c.nb | (c.mode == BEM_LUT ? 32 : 0) | ((cb == 4 ? 0 : 3 - cb) << 6));
    for (int k = 0; k < cb; k++) p[pos + 1 + k] = (uint8_t)((uint32_t)n >> (8 * k));
  }
  pos += 1 + cb;
  if (c.mode == BEM_SIMPLE) { emitPacked(p + pos, q, n, c.nb, lane, a.version); return pos + (int)packedBytes(n, c.nb); }
  if (lane == 0) p[pos] = (uint8_t)(c.nLut + 1);
  pos += 1;
  emitPacked(p + pos, lut + 1, c.nLut, c.nb, lane, a.version); pos += (int)packedBytes(c.nLut, c.nb);   // LUT without the leading 0
  emitPacked(p + pos, rank, n, c.nbIdx, lane, a.version);      pos += (int)packedBytes(n, c.nbIdx);
  return pos;
}

template <class T, bool WRITE>
__global__ void k_tiles(TileArgs a) {
  extern __shared__ __align__(16) uint8_t smemRaw[];
  const int warpsPerCta = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int area = a.mb * a.mb;
  TileSmem<T> s = carveTileSmem<T>(smemRaw + (size_t)warp * tileSmemBytesPerWarp<T>(area), area);
  const T* data = (const T*)a.data;
  const int nBlocks = a.nTx * a.nTy;
  const bool intLossless = !PixelTraits<T>::isFloat && a.maxZErr == 0.5;
  const double scale = a.maxZErr > 0 ? __ddiv_rn(1.0, __dmul_rn(2.0, a.maxZErr)) : 0;

  for (int blk = blockIdx.x * warpsPerCta + warp; blk < nBlocks; blk += gridDim.x * warpsPerCta) {
    if (!WRITE && a.onlyFlagged && a.blockBytes[blk] != 0xffffffffu) continue;
    const int ty = blk / a.nTx, tx = blk - ty * a.nTx;
    const int i0 = ty * a.mb, j0 = tx * a.mb;
    const int h = (i0 + a.mb > a.nRows) ? a.nRows - i0 : a.mb, w = (j0 + a.mb > a.nCols) ? a.nCols - j0 : a.mb;
    const int cells = h * w;
    uint32_t total = 0;
    uint8_t* p = WRITE ? a.out + a.blockOff[blk] : nullptr;

    for (int d = 0; d < a.nDepth; d++) {
      // ---- gather valid pixels in row-major order
      int n = 0;
      for (int base = 0; base < cells; base += 32) {
        const int c = base + lane;
        bool valid = false; T v = 0;
        if (c < cells) {
          const int r = c / w, col = c - r * w;
          const long long k = (long long)(i0 + r) * a.nCols + (j0 + col);
          valid = a.bits ? maskBit(a.bits, k) : true;
          if (valid) v = data[k * a.nDepth + d];
        }
        const unsigned m = __ballot_sync(FULL, valid);
        if (valid) s.buf[n + __popc(m & ((1u << lane) - 1))] = v;
        n += __popc(m);
      }
      __syncwarp();
      if (n == 0 && !WRITE) { total += (uint32_t)a.nDepth; break; }          // Lerc2.cpp:1531-1535

      // ---- min / max / count of equal neighbours
      T lo = n ? s.buf[0] : (T)0, hi = lo;
      int same = 0;
      for (int i = lane; i < n; i += 32) {
        const T v = s.buf[i];
        lo = v < lo ? v : lo; hi = v > hi ? v : hi;
        if (i > 0) same += (v == s.buf[i - 1]) ? 1 : 0;
        else if (a.allValidImage) same += (v == (T)0) ? 1 : 0;               // Lerc2.cpp:1729, :1754
      }
      lo = warpMin(lo); hi = warpMax(hi); same = warpSum(same);
      const double zMin = (double)lo, zMax = (double)hi;
      const bool tryLut = n > 4 && (zMax > __dadd_rn(zMin, __dmul_rn(3.0, a.maxZErr))) && (2 * same > n);
      const bool needQ = needQuantize(a, n, zMin, zMax);
      bool quantDone = false;
      if ((WRITE || tryLut) && needQ) {                                       // Lerc2.h:357-376
        for (int i = lane; i < n; i += 32)
          s.q[i] = intLossless ? (uint32_t)(s.buf[i] - lo) : quantizeOne((double)s.buf[i], zMin, scale);
        quantDone = true;
        __syncwarp();
      }
      (void)quantDone;
      const BlockChoice ca = sizeBlock(a, n, zMin, zMax, (int)sizeof(T), a.dt, tryLut, s.q, s.rank, s.lut, lane);
      int nbAbs = ca.nBytes, nbDiff = nbAbs + 1;

      // ---- depth-delta trial (integer lossless only)                        Lerc2.cpp:1558-1583, :1803-1874
      BlockChoice cd = ca; double dMinD = 0; bool diffOk = false; int32_t dLo = 0;
      if (!PixelTraits<T>::isFloat && a.tryDiff && d > 0 && n > 0) {
        bool overflow = false;
        for (int i = lane; i < n; i += 32) {
          const long long wide = (long long)s.buf[i] - (long long)s.prev[i];
          if (a.checkOverflow && (wide < (long long)INT_MIN || wide > (long long)INT_MAX)) overflow = true;
          s.dbuf[i] = (int32_t)((uint32_t)s.buf[i] - (uint32_t)s.prev[i]);
        }
        __syncwarp();
        diffOk = !__any_sync(FULL, overflow);
        if (diffOk) {
          int32_t l2 = s.dbuf[0], h2 = l2; int same2 = 0;
          for (int i = lane; i < n; i += 32) {
            const int32_t v = s.dbuf[i];
            l2 = v < l2 ? v : l2; h2 = v > h2 ? v : h2;
            same2 += (v == (i > 0 ? s.dbuf[i - 1] : 0)) ? 1 : 0;
          }
          l2 = warpMin(l2); h2 = warpMax(h2); same2 = warpSum(same2);
          dLo = l2; dMinD = (double)l2;
          const double dMaxD = (double)h2;
          const bool tryLutD = n > 4 && (dMaxD > __dadd_rn(dMinD, __dmul_rn(3.0, a.maxZErr))) && (2 * same2 > n);
          if ((WRITE || tryLutD) && needQuantize(a, n, dMinD, dMaxD)) {
            for (int i = lane; i < n; i += 32)
              s.qd[i] = a.maxZErr == 0.5 ? (uint32_t)(s.dbuf[i] - dLo) : quantizeOne((double)s.dbuf[i], dMinD, scale);
            __syncwarp();
          }
          // note: the LUT scratch (rank/lut) is shared with the absolute coding; it is recomputed below if needed
          cd = sizeBlock(a, n, dMinD, dMaxD, 4, DT_Int, tryLutD, s.qd, s.rank, s.lut, lane);
          if (cd.nBytes > 0) nbDiff = cd.nBytes;
        }
      }
      const bool useAbs = (d == 0) || (nbAbs <= nbDiff);                       // absolute coding wins ties (Lerc2.cpp:1640)
      total += (uint32_t)(useAbs ? nbAbs : nbDiff);

      if (WRITE) {
        int wrote;
        if (useAbs) {
          BlockChoice c2 = ca;
          if (ca.mode == BEM_LUT && diffOk) c2 = sizeBlock(a, n, zMin, zMax, (int)sizeof(T), a.dt, tryLut, s.q, s.rank, s.lut, lane);  // redo ranks clobbered by the diff trial
          wrote = emitBlock<T>(a, p, s.buf, n, j0, zMin, c2, false, s.q, s.rank, s.lut, lane);
        } else {
          wrote = emitBlock<int32_t>(a, p, s.dbuf, n, j0, dMinD, cd, true, s.qd, s.rank, s.lut, lane);
        }
        p += wrote;
      }
      if (!PixelTraits<T>::isFloat && a.tryDiff && d < a.nDepth - 1 && n > 0) {
        for (int i = lane; i < n; i += 32) s.prev[i] = s.buf[i];
      }
      __syncwarp();
    }
    if (!WRITE && lane == 0) a.blockBytes[blk] = total;
  }
}

// Count pass for 8-bit lossless rasters where every pixel is valid (the size of the tiling the Huffman modes compete with,
// Lerc2.cpp:261-330): ONE THREAD per 8x8 block position, all DD depths in one sweep over the block's bytes -- minimum, maximum and
// equal-neighbour count of every depth and of every depth-delta block (Lerc2.cpp:1558-1583, :1717-1874), then NumBytesTile
// (Lerc2.h:416-453) for both and the smaller one.  Blocks in which a lookup table has to be tried (more than half of the values
// repeat their predecessor) are left to the general kernel: their entry is 0xffffffff.
template <class T, int DD>
__global__ void __launch_bounds__(128) k_tiles_count8(TileArgs a) {
  const int nBlocks = a.nTx * a.nTy;
  const int blk = blockIdx.x * 128 + threadIdx.x;
  if (blk >= nBlocks) return;
  const uint8_t* data = (const uint8_t*)a.data;
  const int ty = blk / a.nTx, tx = blk - ty * a.nTx;
  const int i0 = ty * 8, j0 = tx * 8;
  const int h = min(8, a.nRows - i0), w = min(8, a.nCols - j0), n = h * w;
  int lo[DD], hi[DD], pv[DD], same[DD], l2[DD], h2[DD], pd[DD], same2[DD];
#pragma unroll
  for (int d = 0; d < DD; d++) { lo[d] = 1 << 20; hi[d] = -(1 << 20); pv[d] = 0; same[d] = 0; l2[d] = 1 << 20; h2[d] = -(1 << 20); pd[d] = 0; same2[d] = 0; }
  bool firstPix = true;
  for (int r = 0; r < h; r++) {
    const uint8_t* row = data + ((size_t)(i0 + r) * a.nCols + j0) * DD;
    uint32_t wds[2 * DD];
    if (w == 8 && ((uintptr_t)row & 7) == 0) {
#pragma unroll
      for (int q = 0; q < DD; q++) { const uint2 x = __ldg((const uint2*)row + q); wds[2 * q] = x.x; wds[2 * q + 1] = x.y; }
    } else {
#pragma unroll
      for (int q = 0; q < 2 * DD; q++) {
        uint32_t x = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) if (q * 4 + b < w * DD) x |= (uint32_t)row[q * 4 + b] << (8 * b);
        wds[q] = x;
      }
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
      if (c < w) {
        int vPrevDepth = 0;
#pragma unroll
        for (int d = 0; d < DD; d++) {
          const int m = c * DD + d;
          const uint32_t b8 = (wds[m >> 2] >> (8 * (m & 3))) & 0xffu;
          const int v = PixelTraits<T>::code == DT_Char ? (int)(int8_t)b8 : (int)b8;
          lo[d] = min(lo[d], v); hi[d] = max(hi[d], v);
          if (!firstPix || a.allValidImage) same[d] += (v == pv[d]) ? 1 : 0;        // Lerc2.cpp:1729, :1754
          pv[d] = v;
          if (d > 0) {
            const int df = v - vPrevDepth;
            l2[d] = min(l2[d], df); h2[d] = max(h2[d], df);
            same2[d] += (df == pd[d]) ? 1 : 0;
            pd[d] = df;
          }
          vPrevDepth = v;
        }
        firstPix = false;
      }
    }
  }
  uint32_t total = 0;
  bool flagged = false;
#pragma unroll
  for (int d = 0; d < DD; d++) {
    const double zMin = (double)lo[d], zMax = (double)hi[d];
    if (n > 4 && (zMax > __dadd_rn(zMin, __dmul_rn(3.0, a.maxZErr))) && (2 * same[d] > n)) flagged = true;
    const BlockChoice ca = sizeBlock(a, n, zMin, zMax, 1, a.dt, false, nullptr, nullptr, nullptr, 0);
    int nbAbs = ca.nBytes, nbDiff = nbAbs + 1;
    if (a.tryDiff && d > 0) {
      const double dMin = (double)l2[d], dMax = (double)h2[d];
      if (n > 4 && (dMax > __dadd_rn(dMin, __dmul_rn(3.0, a.maxZErr))) && (2 * same2[d] > n)) flagged = true;
      const BlockChoice cd = sizeBlock(a, n, dMin, dMax, 4, DT_Int, false, nullptr, nullptr, nullptr, 0);
      if (cd.nBytes > 0) nbDiff = cd.nBytes;
    }
    total += (uint32_t)((d == 0 || nbAbs <= nbDiff) ? nbAbs : nbDiff);
  }
  a.blockBytes[blk] = flagged ? 0xffffffffu : total;
}

template <class T>
static void launchTiles(Context* ctx, const TileArgs& a, bool write) {
  const int area = a.mb * a.mb;
  const int warps = area <= 64 ? 8 : (area <= 256 ? 4 : 1);
  const size_t smem = tileSmemBytesPerWarp<T>(area) * warps;
  const int nBlocks = a.nTx * a.nTy;
  int grid = (nBlocks + warps - 1) / warps;
  if (grid > 148 * 64) grid = 148 * 64;
  if (write) {
    cudaFuncSetAttribute(k_tiles<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LERC_LAUNCH(ctx, (k_tiles<T, true>), grid, warps * 32, smem, a);
  } else {
    cudaFuncSetAttribute(k_tiles<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LERC_LAUNCH(ctx, (k_tiles<T, false>), grid, warps * 32, smem, a);
  }
}

// =================================================================================================
// 8-bit Huffman path                                                            Lerc2.cpp:2311-2468

// nearest valid pixel before k in scan order, or -1
__device__ inline long long prevValidPixel(const uint8_t* __restrict__ bits, long long k) {
  long long i = k - 1;
  while (i >= 0) {
    const unsigned b = bits[i >> 3] & (0xffu << (7 - (i & 7)));   // pixels (i>>3)*8 .. i
    if (b) return (i >> 3) * 8 + 7 - (__ffs(b) - 1);
    i = (i >> 3) * 8 - 1;
  }
  return -1;
}

// predictor of the delta image: left neighbour, else upper, else the last coded value of this depth plane
// (Lerc2.cpp:2336-2341 all-valid, :2362-2371 masked)
template <class T>
__device__ inline T deltaPredictor(const T* __restrict__ data, const uint8_t* __restrict__ bits, long long k, int i, int j, int W, int D, int d) {
  if (j > 0 && (!bits || maskBit(bits, k - 1))) return data[(k - 1) * D + d];
  if (i > 0 && (!bits || maskBit(bits, k - W))) return data[(k - W) * D + d];
  if (!bits) return k > 0 ? data[(k - 1) * D + d] : (T)0;          // only reachable for k == 0
  const long long kp = prevValidPixel(bits, k);
  return kp >= 0 ? data[kp * D + d] : (T)0;
}

template <class T>
__global__ void k_histograms(const T* __restrict__ data, const uint8_t* __restrict__ bits, int H, int W, int D, int* __restrict__ histo /*[512]*/) {
  __shared__ int sh[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int off = PixelTraits<T>::code == DT_Char ? 128 : 0;
  const long long nPix = (long long)H * W;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nPix; k += (long long)gridDim.x * blockDim.x) {
    if (bits && !maskBit(bits, k)) continue;
    const int i = (int)(k / W), j = (int)(k - (long long)i * W);
    for (int d = 0; d < D; d++) {
      const T val = data[k * D + d];
      const T delta = (T)(val - deltaPredictor(data, bits, k, i, j, W, D, d));
      atomicAdd(&sh[off + (int)val], 1);
      atomicAdd(&sh[256 + off + (int)delta], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) if (sh[i]) atomicAdd(&histo[i], sh[i]);
}

struct HuffArgs {
  const void* data; const uint8_t* bits; const uint32_t* chunkBase;   // valid pixels before each 1024-pixel chunk (masked only)
  int H, W, D, delta, numValid;
  uint16_t len[256]; uint32_t code[256];
};

// stream position (symbol index) of (pixel rank r, depth d): depth-planar for the delta mode, pixel-interleaved otherwise
__device__ inline unsigned long long symbolIndex(const HuffArgs& a, unsigned long long r, int d) {
  return a.delta ? (unsigned long long)d * (unsigned long long)a.numValid + r : r * (unsigned long long)a.D + (unsigned long long)d;
}

// Symbols are grouped into segments of 1024 consecutive pixels (per depth plane in delta mode).  Pass 1 sums the
// code lengths of every segment, a scan turns them into bit offsets, pass 2 ORs the codes into a zeroed,
// word-aligned scratch stream (MSB first inside little-endian 32-bit words, Huffman.h:218-255).
template <class T, bool WRITE>
__global__ void __launch_bounds__(256) k_huffman_segments(HuffArgs a, unsigned long long* __restrict__ segBits, const unsigned long long* __restrict__ segOff,
                                   uint32_t* __restrict__ words) {
  // WRITE: a segment's bits are assembled in a per-warp shared-memory window (shared-memory atomics), then stored; only the
  // first and last word of the segment are shared with the neighbouring segments and go through global atomics.
  constexpr int NW = 640;                                  // window words per warp (a 1024-pixel segment at 20 bits/pixel)
  __shared__ uint32_t sWin[WRITE ? 8 * NW : 1];
  // the code table in shared memory (indexing the kernel parameter by a per-thread symbol would put a private copy of it in local memory)
  __shared__ uint32_t sCode[256];
  __shared__ uint16_t sLen[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { sCode[i] = a.code[i]; sLen[i] = a.len[i]; }
  __syncthreads();
  const T* data = (const T*)a.data;
  const int lane = threadIdx.x & 31, warpsPerCta = blockDim.x >> 5, warp = threadIdx.x >> 5;
  const long long nPix = (long long)a.H * a.W;
  const int nChunks = (int)((nPix + 1023) >> 10);
  const long long nSeg = a.delta ? (long long)nChunks * a.D : nChunks;
  const int off = PixelTraits<T>::code == DT_Char ? 128 : 0;
  for (long long seg = (long long)blockIdx.x * warpsPerCta + warp; seg < nSeg; seg += (long long)gridDim.x * warpsPerCta) {
    const int chunk = a.delta ? (int)(seg % nChunks) : (int)seg;
    const int dFirst = a.delta ? (int)(seg / nChunks) : 0, dLast = a.delta ? dFirst + 1 : a.D;
    unsigned long long bitPos = WRITE ? segOff[seg] : 0, total = 0;
    unsigned long long w0 = 0, w1 = 0; bool staged = false;
    uint32_t* win = sWin + (WRITE ? warp * NW : 0);
    if (WRITE) {
      const unsigned long long endBit = segOff[seg + 1];
      w0 = bitPos >> 5; w1 = endBit > bitPos ? ((endBit - 1) >> 5) : w0;
      staged = (w1 - w0 + 1) <= (unsigned long long)NW;
      if (staged) { for (int i = lane; i <= (int)(w1 - w0); i += 32) win[i] = 0; __syncwarp(); }
    }
    for (int step = 0; step < 32; step++) {
      const long long k = (long long)chunk * 1024 + step * 32 + lane;
      const bool valid = k < nPix && (!a.bits || maskBit(a.bits, k));
      unsigned long long myBits = 0;
      const int i = valid ? (int)(k / a.W) : 0, j = valid ? (int)(k - (long long)i * a.W) : 0;
      int syms[4]; int nd = 0;
      if (valid)
        for (int d = dFirst; d < dLast; d++) {
          const T val = data[k * a.D + d];
          const int sym = off + (int)(a.delta ? (T)(val - deltaPredictor(data, a.bits, k, i, j, a.W, a.D, d)) : val);
          myBits += sLen[sym];
          if (nd < 4) syms[nd] = sym;
          nd++;
        }
      // exclusive prefix of myBits over the lanes (stream order inside the step is lane order)
      unsigned long long incl = myBits;
      for (int s = 1; s < 32; s <<= 1) { unsigned long long o = __shfl_up_sync(FULL, incl, s); if (lane >= s) incl += o; }
      const unsigned long long stepTotal = __shfl_sync(FULL, incl, 31);
      if (WRITE && valid) {
        unsigned long long bp = bitPos + incl - myBits;
        int q = 0;
        for (int d = dFirst; d < dLast; d++, q++) {
          int sym;
          if (q < 4) sym = syms[q];
          else { const T val = data[k * a.D + d]; sym = off + (int)(a.delta ? (T)(val - deltaPredictor(data, a.bits, k, i, j, a.W, a.D, d)) : val); }
          const int len = sLen[sym]; const uint32_t code = sCode[sym];
          const unsigned long long wi = bp >> 5; const int used = (int)(bp & 31);
          if (staged) {
            if (32 - used >= len) atomicOr(&win[wi - w0], code << (32 - used - len));
            else { const int spill = len - (32 - used); atomicOr(&win[wi - w0], code >> spill); atomicOr(&win[wi - w0 + 1], code << (32 - spill)); }
          } else {
            if (32 - used >= len) atomicOr(&words[wi], code << (32 - used - len));
            else { const int spill = len - (32 - used); atomicOr(&words[wi], code >> spill); atomicOr(&words[wi + 1], code << (32 - spill)); }
          }
          bp += len;
        }
      }
      bitPos += stepTotal; total += stepTotal;
    }
    if (WRITE && staged) {
      __syncwarp();
      const int nw = (int)(w1 - w0) + 1;
      for (int i = lane; i < nw; i += 32) {
        const uint32_t v = win[i];
        if (i == 0 || i == nw - 1) { if (v) atomicOr(&words[w0 + i], v); }      // shared with the neighbouring segment
        else words[w0 + i] = v;
      }
      __syncwarp();
    }
    if (!WRITE && lane == 0) segBits[seg] = total;
  }
}

// The same two passes for rasters where every pixel is valid, DD = 1 or 3 values per pixel (BASELINE config 4: RGB): a warp takes a
// segment of 1024 pixels, every lane 32 CONSECUTIVE pixels of it (96 bytes: six 16-byte loads, held in registers).  Per bit string
// (one per depth plane in delta mode, one for all depths otherwise) a lane sums its code lengths, one warp scan gives its first
// bit, and the WRITE pass appends its codes in a 64-bit register window that leaves word by word: whole words by plain stores, the
// two words a lane shares with its neighbours by shared-memory atomics (Huffman.h:218-255: MSB first in little-endian words).
template <class T, int DD, bool WRITE>
__global__ void __launch_bounds__(256, 2) k_huffman_runs(HuffArgs a, unsigned long long* __restrict__ segBits, const unsigned long long* __restrict__ segOff,
                                                      uint32_t* __restrict__ words) {
  constexpr int NW = 640;                                  // window words per warp (a 1024-symbol string at 20 bits/symbol)
  __shared__ uint32_t sWin[WRITE ? 8 * NW : 1];
  __shared__ uint32_t sCode[256];
  __shared__ uint8_t sLen[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { sCode[i] = a.code[i]; sLen[i] = (uint8_t)a.len[i]; }
  __syncthreads();
  const uint8_t* data = (const uint8_t*)a.data;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nPix = (long long)a.H * a.W;
  const int nChunks = (int)((nPix + 1023) >> 10);
  const uint32_t flip = PixelTraits<T>::code == DT_Char ? 0x80u : 0u;      // symbol = value (or delta) + 128 for signed char (Lerc2.cpp:2320)
  const bool delta = a.delta != 0;
  uint32_t* win = sWin + (WRITE ? warp * NW : 0);
  for (int chunk = blockIdx.x * 8 + warp; chunk < nChunks; chunk += gridDim.x * 8) {
    const long long k0 = (long long)chunk * 1024 + lane * 32;
    const int nMine = (int)max(0ll, min(32ll, nPix - k0));
    uint32_t v[8 * DD];
    if (nMine == 32) {
      const uint4* src = (const uint4*)(data + k0 * DD);
#pragma unroll
      for (int q = 0; q < 2 * DD; q++) { const uint4 x = __ldg(src + q); v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w; }
    } else {
#pragma unroll
      for (int q = 0; q < 8 * DD; q++) {
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) if ((q * 4 + b) < nMine * DD) w |= (uint32_t)data[k0 * DD + q * 4 + b] << (8 * b);
        v[q] = w;
      }
    }
    const int j0 = nMine > 0 ? (int)(k0 % a.W) : 0;
    // value of byte m of the lane's 32 * DD bytes (m is a compile-time constant wherever this is called)
    auto byteAt = [&](int m) -> uint32_t { return (v[m >> 2] >> (8 * (m & 3))) & 0xffu; };
    // the strings: delta mode DD strings (plane d: segment d * nChunks + chunk), else one (segment chunk)
    const int nStr = delta ? DD : 1;
#pragma unroll
    for (int sI = 0; sI < DD; sI++) {
      if (sI >= nStr) break;
      const long long seg = delta ? (long long)sI * nChunks + chunk : chunk;
      // ---- one sweep over the lane's symbols of string sI; f(sym) is called in stream order
      auto sweep = [&](auto&& f) {
        if (delta) {
          uint32_t prev = 0;
          if (nMine > 0 && k0 > 0) prev = j0 > 0 ? data[(k0 - 1) * DD + sI] : data[(k0 - a.W) * DD + sI];
          int j = j0;
#pragma unroll
          for (int px = 0; px < 32; px++) {
            if (px < nMine) {
              const uint32_t val = byteAt(px * DD + sI);
              if (px > 0 && j == 0) prev = data[(k0 + px - a.W) * DD + sI];      // first pixel of a row: the pixel above
              f(((val - prev) & 0xffu) ^ flip);
              prev = val;
              if (++j == a.W) j = 0;
            }
          }
        } else {
#pragma unroll
          for (int px = 0; px < 32; px++)
            if (px < nMine) {
#pragma unroll
              for (int d = 0; d < DD; d++) f(byteAt(px * DD + d) ^ flip);
            }
        }
      };
      uint32_t myBits = 0;
      sweep([&](uint32_t sym) { myBits += sLen[sym]; });
      uint32_t incl = myBits;
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) { const uint32_t o = __shfl_up_sync(FULL, incl, m); if (lane >= m) incl += o; }
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      if (!WRITE) { if (lane == 0) segBits[seg] = total; continue; }
      const unsigned long long bit0 = segOff[seg], bitEnd = bit0 + total;
      const unsigned long long w0 = bit0 >> 5, w1 = bitEnd > bit0 ? ((bitEnd - 1) >> 5) : w0;
      const int nw = (int)(w1 - w0) + 1;
      const bool staged = nw <= NW;
      if (staged) { for (int i = lane; i < nw; i += 32) win[i] = 0; }
      __syncwarp();
      if (myBits) {
        const unsigned long long myBit = bit0 + (incl - myBits);
        uint32_t* dst = staged ? win + (int)((myBit >> 5) - w0) : words + (myBit >> 5);
        unsigned long long acc = 0;
        int fill = (int)(myBit & 31);                       // leading bits of the first word belong to the predecessor: zero here, OR-ed
        bool first = true;
        sweep([&](uint32_t sym) {
          const int len = sLen[sym];
          acc |= (unsigned long long)sCode[sym] << (64 - fill - len);
          fill += len;
          if (fill >= 32) {
            const uint32_t w = (uint32_t)(acc >> 32);
            if (first) { atomicOr(dst, w); first = false; } else *dst = w;
            dst++; acc <<= 32; fill -= 32;
          }
        });
        if (fill > 0) atomicOr(dst, (uint32_t)(acc >> 32));
      }
      __syncwarp();
      if (staged) {
        for (int i = lane; i < nw; i += 32) {
          const uint32_t x = win[i];
          if (i == 0 || i == nw - 1) { if (x) atomicOr(&words[w0 + i], x); }      // shared with the neighbouring segment
          else words[w0 + i] = x;
        }
        __syncwarp();
      }
    }
  }
}

// raw valid pixels, all depths, in scan order                                    Lerc2.cpp:1343-1364
template <class T>
__global__ void k_one_sweep_gather(const T* __restrict__ data, const uint8_t* __restrict__ bits, const uint32_t* __restrict__ chunkBase,
                                   long long nPix, int nDepth, uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31, warpsPerCta = blockDim.x >> 5;
  const int nChunks = (int)((nPix + 1023) >> 10);
  const size_t len = (size_t)nDepth * sizeof(T);
  for (int c = blockIdx.x * warpsPerCta + (threadIdx.x >> 5); c < nChunks; c += gridDim.x * warpsPerCta) {
    unsigned long long rank = chunkBase[c];
    for (int step = 0; step < 32; step++) {
      const long long k = (long long)c * 1024 + step * 32 + lane;
      const bool valid = k < nPix && maskBit(bits, k);
      const unsigned m = __ballot_sync(FULL, valid);
      if (valid) {
        const uint8_t* src = (const uint8_t*)(data + k * nDepth);
        uint8_t* dst = out + (rank + __popc(m & ((1u << lane) - 1))) * len;
        for (size_t b = 0; b < len; b++) dst[b] = src[b];
      }
      rank += __popc(m);
    }
  }
}

}  // namespace lerc
#include "lerc_encode_fast.cuh"
#include "lerc_encode_tile.cuh"
namespace lerc {

// =================================================================================================
// band orchestration (host)

namespace {

template <class V> bool d2h(Context* ctx, V* hostDst, const void* dSrc, size_t count);

}  // namespace
}  // namespace lerc
#include "lerc_fpl_encode.cuh"
namespace lerc {
namespace {

// Lerc2::TryBitPlaneCompression (Lerc2.cpp:1071-1229): device counts, the reference's decision in double precision on the host
template <class T>
bool tryBitPlane(Context* ctx, const T* dData, const uint8_t* dBitsOrNull, int nRows, int nCols, int nDepth, int numValid, double eps, double& newMaxZErr) {
  newMaxZErr = 0;
  constexpr int maxShift = 8 * (int)sizeof(T), minCnt = 5000;
  if (eps <= 0 || numValid < minCnt) return false;
  const size_t nCounts = (size_t)nDepth * 32 + 1;
  unsigned long long* dCounts = (unsigned long long*)ctx->arena.alloc(nCounts * 8);
  if (!dCounts) return false;
  cudaMemsetAsync(dCounts, 0, nCounts * 8, ctx->stream);
  const long long nPix = (long long)nRows * nCols;
  const int allValid1 = (nDepth == 1 && !dBitsOrNull) ? 1 : 0;
  const dim3 grid((unsigned)std::max<long long>(1, std::min<long long>((nPix + 255) / 256, 148 * 8)), (unsigned)nDepth);
  LERC_LAUNCH(ctx, k_bitplane_counts<T>, grid, 256, 0, dData, dBitsOrNull, nRows, nCols, nDepth, allValid1, dCounts, dCounts + (nCounts - 1));
  std::vector<unsigned long long> h(nCounts);
  if (!d2h(ctx, h.data(), dCounts, nCounts)) return false;
  const unsigned long long cnt = h[nCounts - 1];
  if (cnt < (unsigned long long)minCnt) return false;
  int nCutFound = 0, lastPlaneKept = 0;
  for (int s = maxShift - 1; s >= 0; s--) {
    bool crit = true;
    for (int d = 0; d < nDepth; d++) {
      const double x = (double)h[(size_t)d * 32 + s], n = (double)cnt, m = x / n;
      if (std::fabs(1 - 2 * m) >= eps) crit = false;
    }
    if (crit && nCutFound < 2) {
      if (nCutFound == 0) lastPlaneKept = s;
      if (nCutFound == 1 && s < lastPlaneKept - 1) { lastPlaneKept = s; nCutFound = 0; }
      nCutFound++;
    }
  }
  lastPlaneKept = std::max(0, lastPlaneKept);
  newMaxZErr = (double)((1 << lastPlaneKept) >> 1);
  return true;
}

// Single-pass encoder (lerc_encode_fast.cuh).  Returns true when the band was written (or a definite error is
// in `err`); false when one of its assumptions did not hold and the general encoder must run instead.
template <class T>
bool encodeBandFast(Context* ctx, EncodeBandArgs& a, BandMaskState& ms, uint32_t& bandBytes, ErrCode& err) {
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  using K = typename PixelTraits<T>::Key;
  if (sizeof(T) == 1 || a.nDepth != 1 || a.dValidBytes || a.anyMaskModified || !a.dOut || a.version != 6) return false;
  if (std::getenv("LERC_B200_NO_FAST") || a.maxZErr == 777) return false;     // 777: bit-plane mode (general path)
  double maxZErr = a.maxZErr;
  if (isFlt) { if (!(maxZErr > 0)) return false; }                    // float lossless: FPL / raw decisions stay in the general path
  else maxZErr = std::max(0.5, std::floor(maxZErr));                   // Lerc2.cpp:219
  const long long nPix = (long long)a.nCols * a.nRows;
  const size_t nBits = (size_t)((nPix + 7) >> 3);
  cudaStream_t st = ctx->stream;
  const int nTx = (a.nCols + 7) / 8, nTy = (a.nRows + 7) / 8;
  const long long nBlocks = (long long)nTx * nTy;
  constexpr int TW = EncTile<T>::TW;
  const long long nTiles = (long long)((nTx + TW - 1) / TW) * nTy;      // tiles never wrap a block row
  if (nTiles > 0x7fffffffLL) return false;
  (void)nBlocks;
  const size_t dataStart = (size_t)headerBytes(6) + 4 + 2 * sizeof(T) + 1;
  if (a.outCapacity < dataStart + 1) return false;                     // let the general path report BufferTooSmall exactly

  const size_t nGroups = (size_t)((nTiles + 31) / 32);
  static const bool trace = std::getenv("LERC_B200_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto us = [&]() { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count() * 1e-3; };
  constexpr int MAXSTRIPS = kMaxStrips;
  const size_t stateBytes = sizeof(FastEncResult) + 64 + (size_t)nTiles * 8 + 2 * nGroups * 8;      // result | ticket counters | tile states | group states
  uint8_t* dState = (uint8_t*)ctx->arena.alloc(stateBytes);
  FastEncResult* hRes = (FastEncResult*)ctx->pinnedAlloc(sizeof(FastEncResult));
  unsigned long long* hEnd = (unsigned long long*)ctx->pinnedAlloc(8 * MAXSTRIPS);
  if (!dState || !hRes || !hEnd) return false;
  FastEncResult* dRes = (FastEncResult*)dState;
  cudaMemsetAsync(dState, 0, stateBytes, st);
  // A band that still sits in host memory is coded strip by strip (whole block rows) while its later strips are on their way: host
  // to device copies on one stream, the kernels on the call's stream, and - for a host blob buffer - the finished part of the blob
  // back on a third one, so that the two PCIe directions and the coding overlap.  Look-back state and result block carry over.
  int nStrips = 1, stripAt[kMaxStrips + 1] = {0, nTy};
  if (a.hData) {
    nStrips = stripSchedule((size_t)nPix * sizeof(T), nTy, stripAt);
    if (nStrips > 1 && !ctx->pipeStreams()) { nStrips = 1; stripAt[1] = nTy; }
  }

  FastEncArgs fa;
  fa.data = a.dData; fa.nRows = a.nRows; fa.nCols = a.nCols; fa.nTx = nTx; fa.nTy = nTy; fa.dt = PixelTraits<T>::code;
  fa.maxZErr = maxZErr; fa.scale = 1.0 / (2.0 * maxZErr); fa.maxZErr3 = 3.0 * maxZErr; fa.maxQ = fa.dt <= DT_UShort ? (1u << 15) - 1 : (1u << 30) - 1;         // Lerc2.h:685-703
  fa.intLossless = (!isFlt && maxZErr == 0.5) ? 1 : 0;
  // float data already on a coarser decimal grid (Lerc2.cpp:226-231, :1233-1339): the kernel tests the first row only; if a
  // candidate survives it the general path does the full scan
  fa.nRaise = 0;
  if (isFlt) {
    static const double kErr[9] = {1, 0.5, 0.1, 0.05, 0.01, 0.005, 0.001, 0.0005, 0.0001};
    static const double kFac[9] = {1, 2, 10, 20, 100, 200, 1000, 2000, 10000};
    for (int i = 0; i < 9; i++) if (kErr[i] / 2 > maxZErr) { fa.raiseFac[fa.nRaise] = kFac[i]; fa.nRaise++; }
  }
  uint8_t* blob = a.dOut + a.outOffset;
  fa.stream = blob + dataStart; fa.streamCap = a.outCapacity - dataStart; fa.regionOff = (long long)dataStart - 14;
  fa.tileState = (unsigned long long*)(dState + sizeof(FastEncResult) + 64); fa.res = dRes;
  fa.groupState = fa.tileState + nTiles; fa.groupAcc = fa.groupState + nGroups;   // look-back level 2: groups of 32 tiles
  fa.blob = blob; fa.dataStart = (int)dataStart; fa.nBlobsMore = 0; fa.blobCap = a.outCapacity;
  fa.fillEnd = (a.nBands == 1 && a.fillEnd && a.fillEnd > blob) ? a.fillEnd : nullptr;
  uint8_t* const fillEnd = fa.fillEnd;
  bool outPiped = false;
  {
    // persistent CTAs, tiles taken by ticket (no co-residency assumption); shared memory opt-in and occupancy once per device
    constexpr size_t smem = (size_t)EncTile<T>::SMEM;
    static std::atomic<int> ctasPerSmOf[64];
    const int dv = ctx->device & 63;
    int ctasPerSm = ctasPerSmOf[dv].load(std::memory_order_relaxed);
    if (!ctasPerSm) {
      if (!cudaOk(cudaFuncSetAttribute(k_encode_tile<T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "encode tile smem")) return false;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, k_encode_tile<T, 3>, ENC_THREADS, smem) != cudaSuccess || ctasPerSm < 1) ctasPerSm = 1;
      ctasPerSmOf[dv].store(ctasPerSm, std::memory_order_relaxed);
    }
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int tpr = (nTx + TW - 1) / TW;
    const size_t rowBytes = (size_t)a.nCols * sizeof(T);
    // all host -> device copies first (their stream runs ahead of the kernels)
    if (a.hData) {
      cudaStream_t cin = nStrips > 1 ? ctx->copyIn : st;
      for (int sI = 0; sI < nStrips; sI++) {
        const int br0 = stripAt[sI], br1 = stripAt[sI + 1];
        const size_t r0 = (size_t)br0 * 8, r1 = std::min<size_t>((size_t)br1 * 8, (size_t)a.nRows);
        if (!cudaOk(cudaMemcpyAsync((uint8_t*)const_cast<void*>(a.dData) + r0 * rowBytes, (const uint8_t*)a.hData + r0 * rowBytes, (r1 - r0) * rowBytes,
                                    cudaMemcpyHostToDevice, cin), "H2D strip")) { err = Failed; return true; }
        if (nStrips > 1) cudaEventRecord(ctx->evStrip[0][sI], cin);
      }
      a.hData = nullptr;                                               // (the band is on its way to dData in full, whatever happens below)
    }
    outPiped = nStrips > 1 && a.hOut != nullptr;
    for (int sI = 0; sI < nStrips; sI++) {
      const int br0 = stripAt[sI], br1 = stripAt[sI + 1];
      fa.tileBegin = br0 * tpr; fa.tileEnd = br1 * tpr;
      fa.ticket = (unsigned int*)(dState + sizeof(FastEncResult)) + sI;
      fa.fillEnd = sI == nStrips - 1 ? fillEnd : nullptr;
      fa.hostEnd = outPiped ? hEnd + sI : nullptr;                     // (written by the kernel itself: a small copy would queue behind the blob pieces on the copy engine)
      if (nStrips > 1) cudaStreamWaitEvent(st, ctx->evStrip[0][sI], 0);
      const long long grid = std::min<long long>((long long)(fa.tileEnd - fa.tileBegin), (long long)ctasPerSm * std::max(sms, 1));   // all CTAs resident (the zero fill at the end waits for the last tile)
      { LaunchScope scope_(ctx, "k_encode_tile<T>"); k_encode_tile<T, 3><<<(unsigned)grid, ENC_THREADS, smem, st>>>(fa, FastBatchArgs{}); ctx->kernelLaunches++; }
      if (outPiped) {
        if (sI == nStrips - 1) cudaMemcpyAsync(hRes, dRes, sizeof(FastEncResult), cudaMemcpyDeviceToHost, st);
        cudaEventRecord(ctx->evStrip[1][sI], st);
      }
    }
    fa.fillEnd = fillEnd;
  }
  if (trace) std::fprintf(stderr, "[trace enc] %d strips enqueued at %.1f us\n", nStrips, us());
  if (outPiped) {
    // the finished part of the stream goes to the caller's buffer while the next strips are coded (the prefix follows at the end)
    unsigned long long prev = 0;
    for (int sI = 0; sI < nStrips; sI++) {
      if (!cudaOk(cudaEventSynchronize(ctx->evStrip[1][sI]), "strip sync")) { err = Failed; return true; }
      const unsigned long long end = std::min<unsigned long long>(hEnd[sI], (unsigned long long)(a.outCapacity - dataStart));
      if (end > prev) cudaMemcpyAsync(a.hOut + dataStart + prev, blob + dataStart + prev, (size_t)(end - prev), cudaMemcpyDeviceToHost, ctx->copyOut);
      if (trace) std::fprintf(stderr, "[trace enc] strip %d done at %.1f us, blob bytes %llu..%llu\n", sI, us(), prev, end);
      prev = std::max(prev, end);
    }
  } else {
    if (!cudaOk(cudaMemcpyAsync(hRes, dRes, sizeof(FastEncResult), cudaMemcpyDeviceToHost, st), "D2H fast result")) { err = Failed; return true; }
    if (!cudaOk(cudaStreamSynchronize(st), "sync")) { err = Failed; return true; }
  }
  if (!cudaOk(cudaGetLastError(), "k_encode_tile")) { err = Failed; return true; }
  struct DrainOut { Context* c; bool on; ~DrainOut() { if (on) cudaStreamSynchronize(c->copyOut); } } drainOut{ctx, outPiped};   // (whatever the verdict: no copy into the caller's buffer is left in flight)

  // ---- were the assumptions right?  (On the host: a single warp doing this at the end of the kernel runs cold code at the pace of its
  // instruction fetches, ~8 us measured; here it costs nothing and the prefix follows in stream order.)
  const FastEncResult& r = *hRes;
  if (r.flags & (FASTF_NAN | FASTF_LUT)) return false;
  const K minKey = (K)~r.negMinKey, maxKey = (K)r.maxKey;
  const T lo = fromKeyHost<T>(minKey), hi = fromKeyHost<T>(maxKey);
  const double zMin = (double)lo, zMax = (double)hi;
  if (zMin == zMax) return false;                                        // constant image: no stream at all
  HeaderInfo hd;
  if (isFlt) {
    if ((lo == (T)0 && std::signbit(lo)) || (hi == (T)0 && !std::signbit(hi))) return false;   // sign of a zero extreme depends on scan order (general path)
    bool allInt = !(r.flags & FASTF_NOT_INT);
    const double lim = sizeof(T) == 4 ? 8388608.0 : 9007199254740992.0;
    allInt = allInt && zMin >= -lim && zMin <= lim && zMax >= -lim && zMax <= lim;             // Lerc.cpp:1490-1500
    if (allInt) { if (std::max(0.5, std::floor(maxZErr)) != maxZErr) return false; hd.bIsInt = 1; }
    for (int c = 0; c < fa.nRaise; c++) {                                  // PruneCandidates on row 0 (Lerc2.cpp:1322-1339)
      double m; std::memcpy(&m, &r.raiseMax[c], 8);
      if (!(m / fa.raiseFac[c] > maxZErr / 2)) return false;               // a candidate survived: full scan needed
    }
  }
  const unsigned long long nData = r.totalBytes;
  const size_t oneSweepBytes = sizeof(T) * (size_t)nPix;
  if ((double)nData * 8 < (double)nPix * 1.5 && nData < 4 * oneSweepBytes && (a.nRows > 8 || a.nCols > 8)) return false;   // 16x16 retry (Lerc2.cpp:333-357)
  if (oneSweepBytes <= nData) return false;                              // one sweep raw wins (Lerc2.cpp:364-373)
  const unsigned long long total = dataStart + nData;
  if (total > (unsigned long long)INT_MAX) { err = Failed; return true; }
  bandBytes = (uint32_t)total;
  if (total > a.outCapacity || (r.flags & FASTF_OVERFLOW)) { err = BufferTooSmall; return true; }   // Lerc.cpp:764-765

  // ---- header, mask length, ranges, flag byte; checksum from the kernel's partial sums (Lerc2.cpp:1012-1064)
  hd.version = 6; hd.nRows = a.nRows; hd.nCols = a.nCols; hd.nDepth = 1; hd.dt = PixelTraits<T>::code;
  hd.nBlobsMore = a.nBands - 1 - a.iBand; hd.numValidPixel = (int)nPix; hd.microBlockSize = 8;
  hd.blobSize = (int)total; hd.maxZError = maxZErr; hd.zMin = zMin; hd.zMax = zMax;
  PrefixBytes pb; std::memset(&pb, 0, sizeof pb);
  writeHeader(pb.b, hd);
  size_t p = (size_t)headerBytes(6) + 4;                                  // mask byte count 0
  std::memcpy(pb.b + p, &lo, sizeof(T)); p += sizeof(T);
  std::memcpy(pb.b + p, &hi, sizeof(T)); p += sizeof(T);
  pb.b[p++] = 0;                                                          // not one sweep
  pb.n = (int)p;
  unsigned long long A = 0, D = 0;
  for (int i = 0; i < FAST_SLOTS; i++) { A += r.fletA[i]; D = (D + r.fletD[i]) % 65535ull; }
  fletcherHostPartial(pb.b + 14, 0, (long long)p - 14, A, D);
  hd.checksum = fletcherFinish(A, D, (long long)total - 14);
  std::memcpy(pb.b + 10, &hd.checksum, 4);
  if (trace) std::fprintf(stderr, "[trace enc] verdict at %.1f us\n", us());
  if (outPiped) { std::memcpy(a.hOut, pb.b, (size_t)pb.n); a.hostCopied = true; }        // (the stream's bytes are on their way on copyOut: drained on return)
  else LERC_LAUNCH(ctx, k_write_prefix, 1, 128, 0, blob, pb);
  a.tailFilled = fa.fillEnd != nullptr;

  // validity bookkeeping the band loop expects (Lerc.cpp:659-741): this band is all valid
  ms.numValid = (int)nPix;
  if (a.nBands > 1 && a.iBand < a.nBands - 1) { cudaMemsetAsync(ms.dPrevBits, 0xff, nBits, st); ms.havePrev = true; }
  globalStats().fastPathEncodes++;
  err = cudaOk(cudaGetLastError(), "encodeBandFast") ? Ok : Failed;
  return true;
}

template <class V> bool d2h(Context* ctx, V* hostDst, const void* dSrc, size_t count) {
  if (!cudaOk(cudaMemcpyAsync(hostDst, dSrc, count * sizeof(V), cudaMemcpyDeviceToHost, ctx->stream), "D2H")) return false;
  return cudaOk(cudaStreamSynchronize(ctx->stream), "sync");
}

template <class T>
ErrCode encodeBandT(Context* ctx, EncodeBandArgs& a, BandMaskState& ms, uint32_t& bandBytes) {
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  const long long nPix = (long long)a.nCols * a.nRows;
  const size_t nBits = (size_t)((nPix + 7) >> 3);
  const int nDepth = a.nDepth;
  cudaStream_t st = ctx->stream;
  bandBytes = 0;
  {
    ErrCode fe = Ok;
    const size_t arenaMark = ctx->arena.used, pinnedMark = ctx->pinnedUsed;
    if (encodeBandFast<T>(ctx, a, ms, bandBytes, fe)) return fe;
    if (ctx->arena.retired.empty()) ctx->arena.used = arenaMark;
    ctx->pinnedUsed = pinnedMark;
    bandBytes = 0;
  }
  if (a.hData) {                                                       // the band is still in host memory (the single-pass encoder did not take it)
    if (!cudaOk(cudaMemcpyAsync(const_cast<void*>(a.dData), a.hData, (size_t)nPix * nDepth * sizeof(T), cudaMemcpyHostToDevice, st), "H2D band")) return Failed;
    a.hData = nullptr;
  }

  HeaderInfo hd;
  hd.version = a.version; hd.nRows = a.nRows; hd.nCols = a.nCols; hd.nDepth = nDepth; hd.dt = PixelTraits<T>::code;
  hd.nBlobsMore = a.version >= 6 ? a.nBands - 1 - a.iBand : 0;
  if (a.passNoData) { hd.bPassNoDataValues = 1; hd.noDataVal = a.noDataVal; hd.noDataValOrig = a.noDataOrig; }   // Lerc2.cpp:116-126

  // small device scratch for counters, pulled back through pinned memory
  int* dCounters = (int*)ctx->arena.alloc(64);
  int* hCounters = (int*)ctx->pinnedAlloc(64);
  if (!dCounters || !hCounters) return Failed;

  // ---- 1. validity of this band ---------------------------------------------------------------
  // ms.dBits always holds the bit mask of the band being coded.  dBitsOrNull is what the kernels get:
  // nullptr when every pixel is valid (the reference's "all valid" branches).
  uint8_t* bits = (uint8_t*)ctx->arena.alloc(nBits);
  if (!bits) return Failed;
  int numValid = (int)nPix, maskFlags = 0;
  bool haveBits = false;

  StatsBuffers<T> sb;
  using K = typename PixelTraits<T>::Key;
  sb.minKey = (K*)ctx->arena.alloc(sizeof(K) * nDepth);
  sb.maxKey = (K*)ctx->arena.alloc(sizeof(K) * nDepth);
  sb.zeroIdx = (uint32_t*)ctx->arena.alloc(4 * (size_t)nDepth);
  sb.flags = (int*)ctx->arena.alloc(16);
  sb.ranges = (double*)ctx->arena.alloc(16 * (size_t)nDepth);
  std::vector<double> ranges(2 * (size_t)nDepth);
  if (!sb.minKey || !sb.maxKey || !sb.zeroIdx || !sb.flags || !sb.ranges) return Failed;
  const int pixPerBlock = 8192;
  const dim3 statsGrid((unsigned)((nPix + pixPerBlock - 1) / pixPerBlock), (unsigned)((nDepth + 255) / 256));

  auto runStats = [&](const uint8_t* dBitsOrNull, int& flagsOut) -> bool {
    LERC_LAUNCH(ctx, k_stats_init<T>, (nDepth + 255) / 256, 256, 0, sb, nDepth);
    LERC_LAUNCH(ctx, k_stats<T>, statsGrid, 256, 0, (const T*)a.dData, dBitsOrNull, nPix, nDepth, pixPerBlock, sb);
    LERC_LAUNCH(ctx, k_stats_finish<T>, (nDepth + 255) / 256, 256, 0, (const T*)a.dData, nDepth, sb);
    if (!cudaOk(cudaMemcpyAsync(ranges.data(), sb.ranges, 16 * (size_t)nDepth, cudaMemcpyDeviceToHost, st), "D2H ranges")) return false;
    return d2h(ctx, &flagsOut, sb.flags, 1);
  };
  auto buildMask = [&](const uint8_t* dBytes) -> bool {
    cudaMemsetAsync(dCounters, 0, 8, st);
    launchMaskBuild<T>(ctx, a.dData, dBytes, nPix, nDepth, bits, dCounters);
    if (!d2h(ctx, hCounters, dCounters, 2)) return false;
    numValid = hCounters[0]; maskFlags = hCounters[1];
    haveBits = true;
    return true;
  };

  int statFlags = 0;
  bool statsDone = false;
  if (a.dValidBytes) {
    if (!buildMask(a.dValidBytes)) return Failed;
  } else if (isFlt) {
    if (!runStats(nullptr, statFlags)) return Failed;
    if (statFlags & STATF_NAN) { if (!buildMask(nullptr)) return Failed; }
    else statsDone = true;
  }
  if (maskFlags & MASKF_MIXED_NAN) {
    if (a.version >= 6) return NaNFound;                                     // Lerc.cpp:1481-1484
    // codec versions 2..5 (Lerc::ReplaceNaNValues, Lerc.cpp:901-939): a NaN beside real values of the same pixel becomes -FLT_MAX / -DBL_MAX
    if constexpr (isFlt) {
      T* copy = (T*)ctx->arena.alloc((size_t)nPix * nDepth * sizeof(T));
      if (!copy) return Failed;
      cudaMemcpyAsync(copy, a.dData, (size_t)nPix * nDepth * sizeof(T), cudaMemcpyDeviceToDevice, st);
      const long long nElem = nPix * nDepth;
      LERC_LAUNCH(ctx, k_replace_nan<T>, (int)std::min<long long>((nElem + 255) / 256, 148 * 16), 256, 0, copy, nElem);
      a.dData = copy;
    }
  }
  if (maskFlags & MASKF_MODIFIED) a.anyMaskModified = true;
  if (!haveBits) cudaMemsetAsync(bits, 0xff, nBits, st);

  bool encMask = a.iBand == 0;
  if ((a.nMasks > 1 || a.anyMaskModified) && a.iBand > 0 && ms.havePrev) {  // Lerc.cpp:717-720
    cudaMemsetAsync(dCounters, 0, 4, st);
    launchBitsDiffer(ctx, bits, ms.dPrevBits, nPix, dCounters);
    if (!d2h(ctx, hCounters, dCounters, 1)) return Failed;
    if (hCounters[0]) encMask = true;
  }
  // keep this band's mask for the comparison with the next band (persistent buffer owned by the caller)
  if (a.nBands > 1 && a.iBand < a.nBands - 1) {
    cudaMemcpyAsync(ms.dPrevBits, bits, nBits, cudaMemcpyDeviceToDevice, st);
    ms.havePrev = true;
  }
  ms.dBits = bits; ms.numValid = numValid;
  const uint8_t* dBitsOrNull = (numValid == nPix) ? nullptr : bits;
  hd.numValidPixel = numValid;

  const bool needMask = numValid > 0 && numValid < nPix;
  uint32_t rleBytes = 0;
  uint8_t* dRle = nullptr;
  if (needMask && encMask) {                                                 // Lerc2.cpp:198-203
    dRle = (uint8_t*)ctx->arena.alloc(nBits + nBits / 8000 + 64);
    uint32_t* dSize = (uint32_t*)ctx->arena.alloc(16);
    if (!dRle || !dSize) return Failed;
    launchRleEncode(ctx, bits, (long long)nBits, dRle, dSize);
    if (!d2h(ctx, &rleBytes, dSize, 1)) return Failed;
  }
  const uint32_t headMask = (uint32_t)headerBytes(a.version) + 4 + rleBytes;

  // ---- 2. global decisions -----------------------------------------------------------------------
  double maxZErr = a.maxZErr;
  std::vector<uint8_t> prefix;                      // header .. flags, assembled on the host
  bool oneSweep = false; int imageMode = IEM_Tiling; bool writeTiles = false, writeHuffman = false;
  FplPlan fpl;
  HuffmanTable huff;
  int mbFinal = 8;
  size_t rangesBytes = 0;
  bool constImage = false, allDepthsConst = false;

  if (numValid == 0) {
    if (maxZErr == 777 && (!isFlt || a.version < 6)) { if (isFlt) return Failed; maxZErr = 0; }   // cheat code without enough data: lossless (Lerc2.cpp:210-224)
    if (!isFlt) maxZErr = std::max(0.5, std::floor(maxZErr));
    else if (a.version >= 6) maxZErr = 0;                 // codec version 6 only: the noData / NaN filter resets it for an empty band (Lerc.cpp:1378-1552)
    hd.maxZError = maxZErr; hd.blobSize = (int)headMask;
  } else {
    if (!statsDone && !runStats(dBitsOrNull, statFlags)) return Failed;
    double zMin = ranges[0], zMax = ranges[nDepth];
    for (int m = 1; m < nDepth; m++) { if (ranges[m] < zMin) zMin = ranges[m]; if (zMax < ranges[nDepth + m]) zMax = ranges[nDepth + m]; }
    if (isFlt) {
      bool allInt = !(statFlags & STATF_NOT_INT);
      const double lim = sizeof(T) == 4 ? 8388608.0 : 9007199254740992.0;
      allInt = allInt && zMin >= -lim && zMin <= lim && zMax >= -lim && zMax <= lim;       // Lerc.cpp:1490-1500
      if (a.prefiltered) allInt = a.isAllInt;                                                // decided by prefilterNoData without the noData values
      if (a.version < 6) allInt = false;                                                     // the all-integer rule came with codec version 6 (Lerc.cpp:1490-1502)
      if (allInt) maxZErr = std::max(0.5, std::floor(maxZErr));
      hd.bIsInt = allInt ? 1 : 0;
      if (maxZErr == 777) return Failed;                                                      // the bit-plane "cheat code" is refused for float types (Lerc2.cpp:210-224)
      if (maxZErr > 0) {                                                                      // Lerc2.cpp:226-231
        static const double kErr[9] = {1, 0.5, 0.1, 0.05, 0.01, 0.005, 0.001, 0.0005, 0.0001};
        static const double kFac[9] = {1, 2, 10, 20, 100, 200, 1000, 2000, 10000};
        RaiseArgs ra; double err[9]; ra.n = 0;
        for (int i = 0; i < 9; i++) if (kErr[i] / 2 > maxZErr) { err[ra.n] = kErr[i] / 2; ra.fac[ra.n] = kFac[i]; ra.n++; }
        unsigned long long* dMax = (unsigned long long*)ctx->arena.alloc(9 * 8);
        double hMax[9];
        auto pass = [&](long long r0, long long r1) -> bool {        // rows [r0, r1): returns false when no candidate survives
          if (ra.n == 0) return false;
          cudaMemsetAsync(dMax, 0, 9 * 8, st);
          const long long cnt = (r1 - r0) * a.nCols * nDepth;
          int grid = (int)std::min<long long>((cnt + 255) / 256, 148 * 16);
          LERC_LAUNCH(ctx, k_try_raise<T>, grid, 256, 0, (const T*)a.dData, dBitsOrNull, r0 * a.nCols, r1 * a.nCols, nDepth, ra, dMax);
          if (!d2h(ctx, hMax, dMax, 9)) return false;
          int w = 0;
          for (int c = 0; c < ra.n; c++)
            if (!(hMax[c] / ra.fac[c] > maxZErr / 2)) { err[w] = err[c]; ra.fac[w] = ra.fac[c]; w++; }   // PruneCandidates, Lerc2.cpp:1322-1339
          ra.n = w;
          return w > 0;
        };
        // the first row discards nearly every candidate on real data; only then is the whole band scanned
        if (pass(0, 1) && (a.nRows == 1 || pass(1, a.nRows))) maxZErr = err[0];
      }
    } else {
      if constexpr (!isFlt) {
        if (maxZErr == 777) {                                                                 // bit-plane mode (Lerc2.cpp:210-217, :1071-1229)
          double nz = 0;
          if (!tryBitPlane<T>(ctx, (const T*)a.dData, dBitsOrNull, a.nRows, a.nCols, nDepth, numValid, 0.01, nz)) nz = 0;
          maxZErr = nz;
        }
      }
      maxZErr = std::max(0.5, std::floor(maxZErr));                                          // Lerc2.cpp:219
    }
    hd.maxZError = maxZErr; hd.zMin = zMin; hd.zMax = zMax;
    hd.blobSize = (int)headMask;
    constImage = zMin == zMax;
    if (!constImage) {
      if (a.version >= 4) {                                                                   // Lerc2.cpp:260-281
        rangesBytes = 2 * (size_t)nDepth * sizeof(T);
        if ((size_t)headMask + rangesBytes > (size_t)INT_MAX) return Failed;
        hd.blobSize = (int)(headMask + rangesBytes);
        allDepthsConst = 0 == std::memcmp(ranges.data(), ranges.data() + nDepth, sizeof(double) * nDepth);
      }
    }
  }

  // tile geometry + scratch shared by the count and write passes
  TileArgs ta; std::memset(&ta, 0, sizeof ta);
  uint32_t* dBlockBytes = nullptr; uint32_t* dBlockOff = nullptr;
  size_t nBlocksFinal = 0;
  uint32_t* dChunkBase = nullptr;
  auto ensureChunkBase = [&]() -> bool {
    if (dChunkBase || !dBitsOrNull) return true;
    const int nChunks = (int)((nPix + 1023) >> 10);
    uint32_t* cnt = (uint32_t*)ctx->arena.alloc(4 * (size_t)(nChunks + 1));
    dChunkBase = (uint32_t*)ctx->arena.alloc(4 * (size_t)(nChunks + 1));
    if (!cnt || !dChunkBase) return false;
    cudaMemsetAsync(cnt, 0, 4 * (size_t)(nChunks + 1), st);
    launchChunkValidCounts(ctx, bits, nPix, nChunks, cnt);
    exclusiveScanU32(ctx, cnt, dChunkBase, (size_t)nChunks);
    return true;
  };

  if (numValid > 0 && !constImage && !allDepthsConst) {
    ta.data = a.dData; ta.bits = dBitsOrNull; ta.nRows = a.nRows; ta.nCols = a.nCols; ta.nDepth = nDepth;
    ta.dt = hd.dt; ta.version = a.version; ta.maxZErr = maxZErr;
    ta.maxQ = hd.dt <= DT_UShort ? (1u << 15) - 1 : (1u << 30) - 1;                          // Lerc2.h:685-703
    ta.tryDiff = (a.version >= 5 && !isFlt && nDepth > 1 && maxZErr == 0.5) ? 1 : 0;                           // Lerc2.cpp:1493-1495
    ta.checkOverflow = ((hd.dt == DT_Int || hd.dt == DT_UInt) && (hd.zMax - hd.zMin >= 2147483647.0)) ? 1 : 0;
    ta.allValidImage = numValid == nPix ? 1 : 0;
    auto countPass = [&](int mb, uint32_t*& dLen, uint32_t*& dOff, size_t& nBlocks, long long& total) -> bool {
      ta.mb = mb; ta.nTx = (a.nCols + mb - 1) / mb; ta.nTy = (a.nRows + mb - 1) / mb;
      nBlocks = (size_t)ta.nTx * ta.nTy;
      dLen = (uint32_t*)ctx->arena.alloc(4 * (nBlocks + 1));
      dOff = (uint32_t*)ctx->arena.alloc(4 * (nBlocks + 1));
      if (!dLen || !dOff) return false;
      cudaMemsetAsync(dLen + nBlocks, 0, 4, st);
      ta.blockBytes = dLen; ta.blockOff = nullptr; ta.out = nullptr; ta.onlyFlagged = 0;
      bool counted = false;
      if constexpr (sizeof(T) == 1) {
        if (mb == 8 && !dBitsOrNull && maxZErr == 0.5 && (nDepth == 1 || nDepth == 3) && !std::getenv("LERC_B200_NO_FAST")) {
          const unsigned grid8 = (unsigned)((nBlocks + 127) / 128);
          if (nDepth == 3) LERC_LAUNCH(ctx, (k_tiles_count8<T, 3>), grid8, 128, 0, ta);
          else LERC_LAUNCH(ctx, (k_tiles_count8<T, 1>), grid8, 128, 0, ta);
          ta.onlyFlagged = 1;                                         // the general kernel finishes the blocks that want a lookup table tried
          launchTiles<T>(ctx, ta, false);
          ta.onlyFlagged = 0;
          counted = true;
        }
      }
      if (!counted) launchTiles<T>(ctx, ta, false);
      exclusiveScanU32(ctx, dLen, dOff, nBlocks);
      uint32_t t = 0;
      if (!d2h(ctx, &t, dOff + nBlocks, 1)) return false;
      total = t;
      return true;
    };
    long long nTiling = 0;
    if (!countPass(8, dBlockBytes, dBlockOff, nBlocksFinal, nTiling)) return Failed;
    if (nTiling > INT_MAX) return Failed;
    long long nData = nTiling, nHuff = 0;
    hd.microBlockSize = 8;

    if (hd.tryHuffmanInt()) {                                                                 // Lerc2.cpp:289-304, :2270-2307
      int* dHisto = (int*)ctx->arena.alloc(512 * 4);
      if (!dHisto) return Failed;
      cudaMemsetAsync(dHisto, 0, 512 * 4, st);
      int grid = (int)std::min<long long>((nPix + 255) / 256, 148 * 8);
      LERC_LAUNCH(ctx, k_histograms<T>, grid, 256, 0, (const T*)a.dData, dBitsOrNull, a.nRows, a.nCols, nDepth, dHisto);
      int histo[512];
      if (!d2h(ctx, histo, dHisto, 512)) return Failed;
      HuffmanTable t0, t1; int n0 = 0, n1 = 0;
      if (a.version < 4 || !(t0.buildFromHistogram(histo) && t0.totalBytes(histo, n0))) n0 = 0;    // plain Huffman: codec version >= 4 (Lerc2.cpp:2280)
      if (!(t1.buildFromHistogram(histo + 256) && t1.totalBytes(histo + 256, n1))) n1 = 0;
      if (n0 > 0 || n1 > 0) {
        const bool plain = (n0 > 0 && n1 > 0) ? (n0 <= n1) : (n0 > n1);
        nHuff = plain ? n0 : n1;
        if (nHuff < nTiling) { imageMode = plain ? IEM_Huffman : IEM_DeltaHuffman; huff = plain ? t0 : t1; nData = nHuff; }
      }
    }
    if constexpr (isFlt) {
      if (hd.tryHuffmanFlt()) {                                                               // Lerc2.cpp:305-328: lossless float codec
        if (!planFpl<T>(ctx, (const T*)a.dData, a.dValidBytes, a.nCols, a.nRows, nDepth, fpl)) return Failed;
        nHuff = std::min<long long>(fpl.bytes(), INT_MAX);
        if ((double)nHuff < (double)nTiling * 0.9) { nData = nHuff; imageMode = IEM_DeltaDeltaHuffman; }   // at least 10 % better than tiling
      }
    }
    // 16x16 retry when the bit rate is tiny (Lerc2.cpp:333-357)
    const size_t oneSweepBytes = sizeof(T) * (size_t)nDepth * (size_t)numValid;
    if (((size_t)nTiling * 8 < (size_t)nPix * nDepth * 1.5) && ((size_t)nTiling < 4 * oneSweepBytes) &&
        (nHuff == 0 || (size_t)nTiling < (size_t)2 * (size_t)nHuff) && (a.nRows > 8 || a.nCols > 8)) {
      uint32_t* dLen2; uint32_t* dOff2; size_t nBlocks2; long long n2 = 0;
      if (!countPass(16, dLen2, dOff2, nBlocks2, n2)) return Failed;
      if (n2 <= nData) { nData = n2; imageMode = IEM_Tiling; hd.microBlockSize = 16; dBlockBytes = dLen2; dBlockOff = dOff2; nBlocksFinal = nBlocks2; }
    }
    mbFinal = hd.microBlockSize;
    if (hd.tryHuffmanInt() || hd.tryHuffmanFlt()) nData += 1;                                // image-mode flag byte
    size_t total = (size_t)hd.blobSize;
    if (oneSweepBytes <= (size_t)nData) { oneSweep = true; total += 1 + oneSweepBytes; }       // Lerc2.cpp:364-373
    else total += 1 + (size_t)nData;
    if (total > (size_t)INT_MAX) return Failed;
    hd.blobSize = (int)total;
    writeHuffman = !oneSweep && imageMode != IEM_Tiling;
    writeTiles = !oneSweep && !writeHuffman;
  }

  bandBytes = (uint32_t)hd.blobSize;
  if (!a.dOut) return Ok;                                                                     // size-only call
  if ((size_t)bandBytes > a.outCapacity) return BufferTooSmall;                               // Lerc.cpp:764-765

  // ---- 3. write ---------------------------------------------------------------------------------
  uint8_t* blob = a.dOut + a.outOffset;
  const size_t hb = (size_t)headerBytes(a.version);
  // (a) header + mask length field
  uint8_t* hHead = (uint8_t*)ctx->pinnedAlloc(hb + 4);
  if (!hHead) return Failed;
  writeHeader(hHead, hd);
  const int32_t nm = (int32_t)rleBytes;
  std::memcpy(hHead + hb, &nm, 4);
  cudaMemcpyAsync(blob, hHead, hb + 4, cudaMemcpyHostToDevice, st);
  size_t pos = hb + 4;
  if (rleBytes) { cudaMemcpyAsync(blob + pos, dRle, rleBytes, cudaMemcpyDeviceToDevice, st); pos += rleBytes; }
  // (b) ranges + flag bytes [+ Huffman table]
  if (numValid > 0 && !constImage) {
    std::vector<uint8_t> tail(rangesBytes + 2 + 4096, 0);
    size_t tp = 0;
    if (a.version >= 4) {
      for (int m = 0; m < nDepth; m++) { T v = (T)ranges[m]; std::memcpy(tail.data() + tp, &v, sizeof(T)); tp += sizeof(T); }
      for (int m = 0; m < nDepth; m++) { T v = (T)ranges[nDepth + m]; std::memcpy(tail.data() + tp, &v, sizeof(T)); tp += sizeof(T); }
    }
    size_t huffTableBytes = 0;
    if (!allDepthsConst) {
      tail[tp++] = oneSweep ? 1 : 0;
      if (!oneSweep && (hd.tryHuffmanInt() || hd.tryHuffmanFlt())) tail[tp++] = (uint8_t)imageMode;
      if (writeHuffman && !isFlt) { huffTableBytes = huff.write(tail.data() + tp, a.version); if (!huffTableBytes) return Failed; tp += huffTableBytes; }
    }
    uint8_t* hTail = (uint8_t*)ctx->pinnedAlloc(tp);
    if (!hTail) return Failed;
    std::memcpy(hTail, tail.data(), tp);
    cudaMemcpyAsync(blob + pos, hTail, tp, cudaMemcpyHostToDevice, st);
    pos += tp;

    if (!allDepthsConst) {
      if (oneSweep) {
        const size_t len = sizeof(T) * (size_t)nDepth;
        if (!dBitsOrNull) cudaMemcpyAsync(blob + pos, a.dData, len * (size_t)nPix, cudaMemcpyDeviceToDevice, st);
        else {
          if (!ensureChunkBase()) return Failed;
          const int nChunks = (int)((nPix + 1023) >> 10);
          LERC_LAUNCH(ctx, k_one_sweep_gather<T>, std::min((nChunks + 7) / 8, 148 * 8), 256, 0, (const T*)a.dData, bits, dChunkBase, nPix, nDepth, blob + pos);
        }
        pos += len * (size_t)numValid;
      } else if (writeHuffman) {
        if constexpr (isFlt) {                                                                // Lerc2.cpp:437-449
          size_t w = 0;
          if (imageMode != IEM_DeltaDeltaHuffman || !writeFpl(ctx, fpl, blob + pos, w)) return Failed;
          pos += w;
        } else if constexpr (sizeof(T) == 1) {
          if (!ensureChunkBase()) return Failed;
          HuffArgs ha;
          ha.data = a.dData; ha.bits = dBitsOrNull; ha.chunkBase = dChunkBase; ha.H = a.nRows; ha.W = a.nCols; ha.D = nDepth;
          ha.delta = imageMode == IEM_DeltaHuffman ? 1 : 0; ha.numValid = numValid;
          std::memcpy(ha.len, huff.len, sizeof ha.len); std::memcpy(ha.code, huff.code, sizeof ha.code);
          const int nChunks = (int)((nPix + 1023) >> 10);
          const size_t nSeg = ha.delta ? (size_t)nChunks * nDepth : (size_t)nChunks;
          unsigned long long* dSegBits = (unsigned long long*)ctx->arena.alloc(8 * (nSeg + 1));
          unsigned long long* dSegOff = (unsigned long long*)ctx->arena.alloc(8 * (nSeg + 1));
          if (!dSegBits || !dSegOff) return Failed;
          cudaMemsetAsync(dSegBits + nSeg, 0, 8, st);
          int grid = (int)std::min<size_t>((nSeg + 7) / 8, 148 * 16);
          // all pixels valid, 1 or 3 values per pixel, 16-byte aligned: the register-window kernels
          const int runsD = (!dBitsOrNull && ((uintptr_t)a.dData & 15) == 0 && nPix < (1ll << 31)) ? (nDepth == 1 ? 1 : (nDepth == 3 ? 3 : 0)) : 0;
          const int gridRuns = (int)std::min<size_t>(((size_t)nChunks + 7) / 8, 148 * 16);
          if (runsD == 3) LERC_LAUNCH(ctx, (k_huffman_runs<T, 3, false>), gridRuns, 256, 0, ha, dSegBits, nullptr, nullptr);
          else if (runsD == 1) LERC_LAUNCH(ctx, (k_huffman_runs<T, 1, false>), gridRuns, 256, 0, ha, dSegBits, nullptr, nullptr);
          else LERC_LAUNCH(ctx, (k_huffman_segments<T, false>), grid, 256, 0, ha, dSegBits, nullptr, nullptr);
          exclusiveScanU64(ctx, dSegBits, dSegOff, nSeg);
          // The masked delta mode codes valid pixels only, so segment order == stream order holds in both modes
          // as long as the delta planes are laid out depth after depth, which the segment numbering does.
          const size_t dataBytes = (size_t)hd.blobSize - pos;       // bit stream words + the read-ahead word
          uint32_t* dWords = (uint32_t*)ctx->arena.alloc(dataBytes + 16);
          if (!dWords) return Failed;
          cudaMemsetAsync(dWords, 0, dataBytes + 16, st);
          if (runsD == 3) LERC_LAUNCH(ctx, (k_huffman_runs<T, 3, true>), gridRuns, 256, 0, ha, nullptr, dSegOff, dWords);
          else if (runsD == 1) LERC_LAUNCH(ctx, (k_huffman_runs<T, 1, true>), gridRuns, 256, 0, ha, nullptr, dSegOff, dWords);
          else LERC_LAUNCH(ctx, (k_huffman_segments<T, true>), grid, 256, 0, ha, nullptr, dSegOff, dWords);
          cudaMemcpyAsync(blob + pos, dWords, dataBytes, cudaMemcpyDeviceToDevice, st);
          pos += dataBytes;
        } else return Failed;
      } else if (writeTiles) {
        ta.mb = mbFinal; ta.nTx = (a.nCols + mbFinal - 1) / mbFinal; ta.nTy = (a.nRows + mbFinal - 1) / mbFinal;
        ta.blockBytes = nullptr; ta.blockOff = dBlockOff; ta.out = blob + pos;
        launchTiles<T>(ctx, ta, true);
        pos = (size_t)hd.blobSize;
      }
    }
  }
  if (pos != (size_t)hd.blobSize) {
    if (std::getenv("LERC_B200_VERBOSE")) std::fprintf(stderr, "[lerc_b200] internal size mismatch: wrote %zu, planned %d\n", pos, hd.blobSize);
    return Failed;
  }
  // (c) checksum over [14, blobSize) stored at byte 10                        Lerc2.cpp:1012-1030
  if (a.version >= 3) {
    unsigned long long* dAcc = (unsigned long long*)ctx->arena.alloc(16);
    if (!dAcc) return Failed;
    cudaMemsetAsync(dAcc, 0, 16, st);
    launchFletcher(ctx, blob + 14, (long long)hd.blobSize - 14, dAcc, blob + 10, 0, nullptr);
  }
  return cudaOk(cudaGetLastError(), "encodeBand") ? Ok : Failed;
}

}  // namespace

ErrCode encodeBand(Context* ctx, EncodeBandArgs& a, BandMaskState& ms, uint32_t& bandBytes) {
  switch (a.dt) {
    case DT_Char:   return encodeBandT<int8_t>(ctx, a, ms, bandBytes);
    case DT_Byte:   return encodeBandT<uint8_t>(ctx, a, ms, bandBytes);
    case DT_Short:  return encodeBandT<int16_t>(ctx, a, ms, bandBytes);
    case DT_UShort: return encodeBandT<uint16_t>(ctx, a, ms, bandBytes);
    case DT_Int:    return encodeBandT<int32_t>(ctx, a, ms, bandBytes);
    case DT_UInt:   return encodeBandT<uint32_t>(ctx, a, ms, bandBytes);
    case DT_Float:  return encodeBandT<float>(ctx, a, ms, bandBytes);
    case DT_Double: return encodeBandT<double>(ctx, a, ms, bandBytes);
    default: return WrongParam;
  }
}

}  // namespace lerc

// =================================================================================================
// bands with a caller-supplied noData value (the _4D calls): Lerc::FilterNoDataAndNaN (Lerc.cpp:1378-1552, float types) and
// Lerc::FilterNoData (Lerc.cpp:1241-1374, integer types), on private device copies of the band and its byte mask
namespace lerc {
namespace {

enum { NDF_MODIFIED = 1, NDF_LEFT = 2, NDF_NOT_INT = 4, NDF_ANY_VALID = 8 };
struct NoDataScan { unsigned long long minKey, negMaxKey; unsigned int flags, pad; };   // minKey starts at ~0, negMaxKey (= ~maxKey) at ~0

template <class T>
__global__ void k_nodata_scan(T* __restrict__ data, uint8_t* __restrict__ maskBytes, long long nPix, int nDepth, double noDataD, NoDataScan* __restrict__ out) {
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  using K = typename PixelTraits<T>::Key;
  const T origNoData = (T)noDataD;
  K kmin = keyMaxValue<K>(), kmax = 0;
  unsigned int flags = 0;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nPix; k += (long long)gridDim.x * blockDim.x) {
    if (!maskBytes[k]) continue;
    T* px = data + k * nDepth;
    int bad = 0;
    for (int m = 0; m < nDepth; m++) {
      const T z = px[m];
      if (isFlt && isNaNVal(z)) { bad++; px[m] = nDepth > 1 ? origNoData : (T)0; }        // Lerc.cpp:1436-1444
      else if (z == origNoData) bad++;
      else {
        const K key = toKey(z);
        kmin = key < kmin ? key : kmin; kmax = key > kmax ? key : kmax;
        flags |= NDF_ANY_VALID;
        if (isFlt && !(z == (T)floor((double)z + 0.5))) flags |= NDF_NOT_INT;             // Lerc.h:248
      }
    }
    if (bad == nDepth) { maskBytes[k] = 0; flags |= NDF_MODIFIED; }
    else if (bad > 0) flags |= NDF_LEFT;
  }
  kmin = warpMin(kmin); kmax = warpMax(kmax);
  flags = __reduce_or_sync(FULL, flags);
  if ((threadIdx.x & 31) == 0 && flags) {
    if (flags & NDF_ANY_VALID) { atomicMin(&out->minKey, (unsigned long long)kmin); atomicMin(&out->negMaxKey, ~(unsigned long long)kmax); }
    atomicOr(&out->flags, flags);
  }
}

template <class T>
__global__ void k_nodata_remap(T* __restrict__ data, const uint8_t* __restrict__ maskBytes, long long nPix, int nDepth, double fromD, double toD) {
  const T from = (T)fromD, to = (T)toD;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nPix; k += (long long)gridDim.x * blockDim.x) {
    if (!maskBytes[k]) continue;
    T* px = data + k * nDepth;
    for (int m = 0; m < nDepth; m++) if (px[m] == from) px[m] = to;
  }
}

template <class T> bool isIntegralHost(T z) { return z == (T)std::floor((double)z + 0.5); }

// Lerc::FindNewNoDataBelowValidMin (Lerc.cpp:1556-1618)
template <class T>
bool findNewNoData(double minVal, double maxZErr, bool allInt, double lowIntLimit, T& out) {
  std::vector<T> cand;
  if (allInt) {
    for (double dist : {4 * maxZErr, 1.0, 10.0, 100.0, 1000.0, 10000.0}) cand.push_back((T)(minVal - dist));
    cand.push_back((T)(minVal > 0 ? std::floor(minVal / 2) : minVal * 2));
  } else {
    for (double dist : {4 * maxZErr, 0.0001, 0.001, 0.01, 0.1, 1.0, 10.0, 100.0, 1000.0, 10000.0}) cand.push_back((T)(minVal - dist));
    cand.push_back((T)(minVal > 0 ? minVal / 2 : minVal * 2));
  }
  std::sort(cand.begin(), cand.end(), [](T x, T y) { return x > y; });
  const T lowest = (T)(sizeof(T) == 4 ? -FLT_MAX : -DBL_MAX);
  for (T v : cand)
    if (allInt ? (v > (T)lowIntLimit && v < (T)(minVal - 2 * maxZErr) && isIntegralHost(v)) : (v > lowest && v < (T)(minVal - 2 * maxZErr))) { out = v; return true; }
  return false;
}

template <class T>
ErrCode prefilterNoDataT(Context* ctx, void* dData, uint8_t* dMaskBytes, long long nPix, int nDepth, double maxZErr, double noDataOrigD, NoDataVerdict& out) {
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  using K = typename PixelTraits<T>::Key;
  out.maxZErr = maxZErr; out.noDataVal = noDataOrigD; out.maskModified = out.needNoData = out.isAllInt = false;
  double tLo = 0, tHi = 0;
  if (isFlt) { if (sizeof(T) == 4 && (noDataOrigD < -FLT_MAX || noDataOrigD > FLT_MAX)) return WrongParam; }
  else {
    static const double kLo[6] = {-128, 0, -32768, 0, -2147483648.0, 0}, kHi[6] = {127, 255, 32767, 65535, 2147483647.0, 4294967295.0};   // GetTypeRange
    tLo = kLo[PixelTraits<T>::code]; tHi = kHi[PixelTraits<T>::code];
    if (noDataOrigD < tLo || noDataOrigD > tHi) return WrongParam;
  }
  const T origNoData = (T)noDataOrigD;
  cudaStream_t st = ctx->stream;
  NoDataScan* dScan = (NoDataScan*)ctx->arena.alloc(sizeof(NoDataScan));
  NoDataScan* hScan = (NoDataScan*)ctx->pinnedAlloc(sizeof(NoDataScan));
  if (!dScan || !hScan) return Failed;
  hScan->minKey = ~0ull; hScan->negMaxKey = ~0ull; hScan->flags = 0; hScan->pad = 0;
  cudaMemcpyAsync(dScan, hScan, sizeof(NoDataScan), cudaMemcpyHostToDevice, st);
  const int grid = (int)std::max<long long>(1, std::min<long long>((nPix + 255) / 256, 148 * 16));
  LERC_LAUNCH(ctx, k_nodata_scan<T>, grid, 256, 0, (T*)dData, dMaskBytes, nPix, nDepth, noDataOrigD, dScan);
  if (!cudaOk(cudaMemcpyAsync(hScan, dScan, sizeof(NoDataScan), cudaMemcpyDeviceToHost, st), "D2H noData scan") || !cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
  const unsigned fl = hScan->flags;
  out.maskModified = (fl & NDF_MODIFIED) != 0;
  const bool left = (fl & NDF_LEFT) != 0;
  if (!(fl & NDF_ANY_VALID)) { out.maxZErr = isFlt ? 0 : 0.5; return Ok; }                         // no valid data in this band
  const double minVal = (double)fromKeyHost<T>((K)hScan->minKey), maxVal = (double)fromKeyHost<T>((K)~hScan->negMaxKey);
  auto remap = [&](T to) {
    LERC_LAUNCH(ctx, k_nodata_remap<T>, grid, 256, 0, (T*)dData, (const uint8_t*)dMaskBytes, nPix, nDepth, noDataOrigD, (double)to);
    out.noDataVal = (double)to;
  };
  if constexpr (isFlt) {
    out.needNoData = left;
    const double lowIntLimit = sizeof(T) == 4 ? -8388608.0 : -9007199254740992.0, highIntLimit = -lowIntLimit;
    bool allInt = !(fl & NDF_NOT_INT);
    double mzL = maxZErr;
    if (allInt) {
      allInt = minVal >= lowIntLimit && minVal <= highIntLimit && maxVal >= lowIntLimit && maxVal <= highIntLimit;
      if (left) allInt = allInt && isIntegralHost(origNoData) && origNoData >= lowIntLimit && origNoData <= highIntLimit;
      if (allInt) mzL = std::max(0.5, std::floor(maxZErr));
    }
    out.isAllInt = allInt;
    if (mzL == 0) return Ok;
    const double dist = allInt ? std::floor(mzL) : 2 * mzL;
    if (origNoData >= minVal - dist && origNoData <= maxVal + dist) { out.maxZErr = allInt ? 0.5 : 0; return Ok; }   // fall back to lossless
    if (left) {
      T to = origNoData;
      if (findNewNoData<T>(minVal, mzL, allInt, lowIntLimit, to)) { if (to != origNoData) remap(to); }
      else if ((double)origNoData >= minVal) mzL = allInt ? 0.5 : 0;
    }
    out.maxZErr = mzL;
  } else {
    out.needNoData = left;
    double mzL = std::max(0.5, std::floor(maxZErr));
    const double dist = std::floor(mzL);
    if (origNoData >= minVal - dist && origNoData <= maxVal + dist) { out.maxZErr = 0.5; return Ok; }
    if (left) {
      const double minDist = std::floor(mzL) + 1;
      double remapVal = minVal - minDist;
      T to = origNoData;
      if (remapVal >= tLo) to = (T)remapVal;
      else {
        mzL = 0.5;
        remapVal = minVal - 1;
        if (remapVal >= tLo) to = (T)remapVal;
        else { remapVal = maxVal + 1; if (remapVal <= tHi && remapVal < origNoData) to = (T)remapVal; }
      }
      if (to != origNoData) remap(to);
    }
    out.maxZErr = mzL;
  }
  return cudaOk(cudaGetLastError(), "prefilterNoData") ? Ok : Failed;
}

}  // namespace

ErrCode prefilterNoData(Context* ctx, int dt, void* dData, uint8_t* dMaskBytes, long long nPix, int nDepth, double maxZErr, double noDataOrig,
                        NoDataVerdict& out) {
  switch (dt) {
    case DT_Char:   return prefilterNoDataT<int8_t>(ctx, dData, dMaskBytes, nPix, nDepth, maxZErr, noDataOrig, out);
    case DT_Byte:   return prefilterNoDataT<uint8_t>(ctx, dData, dMaskBytes, nPix, nDepth, maxZErr, noDataOrig, out);
    case DT_Short:  return prefilterNoDataT<int16_t>(ctx, dData, dMaskBytes, nPix, nDepth, maxZErr, noDataOrig, out);
    case DT_UShort: return prefilterNoDataT<uint16_t>(ctx, dData, dMaskBytes, nPix, nDepth, maxZErr, noDataOrig, out);
    case DT_Int:    return prefilterNoDataT<int32_t>(ctx, dData, dMaskBytes, nPix, nDepth, maxZErr, noDataOrig, out);
    case DT_UInt:   return prefilterNoDataT<uint32_t>(ctx, dData, dMaskBytes, nPix, nDepth, maxZErr, noDataOrig, out);
    case DT_Float:  return prefilterNoDataT<float>(ctx, dData, dMaskBytes, nPix, nDepth, maxZErr, noDataOrig, out);
    case DT_Double: return prefilterNoDataT<double>(ctx, dData, dMaskBytes, nPix, nDepth, maxZErr, noDataOrig, out);
    default: return WrongParam;
  }
}

}  // namespace lerc

#include "lerc_tiles_encode.cuh"
