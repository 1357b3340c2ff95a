// lerc_encode_fast.cuh -- single-pass fused Lerc2 band encoder for the common raster shape:
// every pixel valid, nDepth == 1, 8x8 micro-blocks, 16/32/64-bit pixel types (included by lerc_encode.cu).
//
// One kernel reads the raster ONCE and writes the finished micro-block stream ONCE:
//   per CTA tile = TILE_BLOCKS consecutive micro-blocks in stream order (Lerc2.cpp:1507-1521)
//     8 lanes per micro-block, one lane per block row (8 pixels in registers, 128-bit global loads)
//     block min/max by 3 xor-shuffles in order-preserving key space   (GetValidDataAndStats, Lerc2.cpp:1717-1799)
//     block coding choice + byte length                               (NumBytesTile, Lerc2.h:416-453)
//     CTA exclusive scan of the lengths, decoupled look-back across tiles for the global byte offset
//     fp64 quantisation without contraction                           (Quantize, Lerc2.h:357-376)
//     each block row packs to w*numBits bits in registers (funnel shifts), OR-ed into a shared-memory staging
//     image of the tile's output bytes                                (WriteTile Lerc2.cpp:1949-2021, BitStuffer2.cpp:35-75, :432-472)
//     staging -> HBM with 128-bit stores aligned to the global address, Fletcher-32 partial sums of exactly
//     those bytes                                                     (Lerc2.cpp:1037-1064)
//   image-global facts (min/max, NaN, all-integer, LUT candidates) fall out of the same pass.
// The kernel is speculative about the image-global decisions the reference takes before block coding
// (Lerc2.cpp:179-381): it assumes "no NaN, maxZError as given, 8x8 tiling wins, no LUT block".  The host checks
// the facts afterwards and re-runs the general multi-pass encoder (encodeBandT) when any assumption failed, so
// the bytes are always those of the reference.
#pragma once
#include <type_traits>

namespace lerc {

enum { FASTF_NAN = 1, FASTF_NOT_INT = 2, FASTF_LUT = 4, FASTF_OVERFLOW = 8 };
constexpr int FAST_SLOTS = 32;

struct FastEncResult {                 // device, zero-initialised per call
  unsigned long long totalBytes;       // length of the micro-block stream
  unsigned long long negMinKey;        // ~min key (so that zero-init works with atomicMax)
  unsigned long long maxKey;
  unsigned int flags, ticket;
  unsigned long long fletA[FAST_SLOTS], fletD[FAST_SLOTS];
  unsigned long long raiseMax[9];      // row 0: max |round(x * fac) - x * fac| per TryRaiseMaxZError candidate, as double bits
  unsigned int totalReady, pad;        // totalBytes is final
};

struct FastEncArgs {
  const void* data; int nRows, nCols, nTx, nTy, dt;
  double maxZErr, scale, maxZErr3; uint32_t maxQ;   // scale = 1 / (2 * maxZErr) (Lerc2.h:339), maxZErr3 = 3 * maxZErr (Lerc2.cpp:1794), both rounded on the host as the reference does
  int intLossless;                     // integer type && maxZErr == 0.5
  uint8_t* stream;                     // where the micro-block stream starts (blob + dataStart)
  unsigned long long streamCap;        // bytes available from `stream`
  long long regionOff;                 // checksum-region offset of stream[0] (= dataStart - 14)
  unsigned long long* tileState;       // [nTiles], zero-initialised
  FastEncResult* res;
  unsigned long long* groupState;      // [ceil(nTiles / 32)], zero-initialised: aggregates of 32 consecutive tiles (two-round look-back); nullptr = plain chain
  unsigned long long* groupAcc;        // [ceil(nTiles / 32)], zero-initialised: atomic accumulators of the groups (lerc_lookback.cuh)
  // k_encode_tile only: the tiles [tileBegin, tileEnd) of this launch (a band may be coded strip by strip while its rows arrive from the
  // host; look-back state and result block carry over), their ticket counter (zero-initialised, one per launch);
  // row-0 test of TryRaiseMaxZError (Lerc2.cpp:1233-1339), zero fill behind the blob
  unsigned long long* hostEnd;     // mapped host word or nullptr: where the stream ends behind tile tileEnd - 1 (strip pipelining of host-resident calls)
  int tileBegin, tileEnd; unsigned int* ticket;
  double raiseFac[9]; int nRaise;
  uint8_t* blob; int dataStart, nBlobsMore;            // where the band blob starts; header field
  unsigned long long blobCap;                          // bytes available from `blob`
  uint8_t* fillEnd;                                    // zero-fill [blob + size, fillEnd) (Lerc.cpp:374) or nullptr: the caller does it
};

// ---- tile batch (k_encode_fused<T, MINB, true>, lerc_tiles_encode.cuh): the raster is cut into imgRows x imgCols images, every
// image becomes its own blob.  nTx / nTy / nRows / nCols of FastEncArgs then describe a FULL image; every image owns segPerImg
// consecutive entries of tileState (edge images leave some of them empty).
struct TileEncResult;
struct FastBatchArgs {
  int imgCols, imgRows, nImgX, nImgY, rasterCols, rasterRows, segPerImg, dataStart;
  long long pitch;                     // elements per raster row
  unsigned long long* imgState;        // [nImg], zero-initialised: look-back over whole blobs
  TileEncResult* imgRes;               // [nImg], zero-initialised
  uint8_t* out; unsigned long long outCap;
  // k_encode_tile<T, MINB, true>: tiles of TW consecutive micro-blocks per image
  int tilesPerImg;
  uint32_t* tileLen;                   // [nImg * tilesPerImg] byte count of every tile (the look-back words lose it when they turn into prefixes)
};
struct FastNoBatch {                   // what the one-image kernel gets instead: nothing (the names fold to constants)
  static constexpr int imgCols = 0, imgRows = 0, nImgX = 0, nImgY = 0, rasterCols = 0, rasterRows = 0, segPerImg = 1, dataStart = 0;
  static constexpr long long pitch = 0;
  static constexpr unsigned long long* imgState = nullptr;
  static constexpr TileEncResult* imgRes = nullptr;
  static constexpr uint8_t* out = nullptr; static constexpr unsigned long long outCap = 0;
};

struct TileEncResult {                 // per image of a tile batch
  unsigned long long negMinKey, maxKey, fletA, fletD, streamBytes;
  unsigned int flags, pad;
};

// ---- helpers -----------------------------------------------------------------------------------
template <class K> __device__ __forceinline__ K shflXorK(K v, int m);
template <> __device__ __forceinline__ uint32_t shflXorK<uint32_t>(uint32_t v, int m) { return __shfl_xor_sync(FULL, v, m); }
template <> __device__ __forceinline__ unsigned long long shflXorK<unsigned long long>(unsigned long long v, int m) { return __shfl_xor_sync(FULL, v, m); }

// OR the bit string R (little-endian words, nbits long) into the zeroed staging words at bit offset bitOff.
template <int NW>
__device__ __forceinline__ void orBits(uint32_t* __restrict__ stage, uint32_t bitOff, const uint32_t (&R)[NW], int nbits) {
  const uint32_t w0 = bitOff >> 5, sh = bitOff & 31;
  const int nw = (int)((sh + (uint32_t)nbits + 31) >> 5);
  uint32_t prev = 0;
#pragma unroll
  for (int j = 0; j <= NW; j++) {
    const uint32_t cur = j < NW ? R[j] : 0u;
    const uint32_t o = __funnelshift_l(prev, cur, sh);
    if (j < nw && o) atomicOr(&stage[w0 + j], o);
    prev = cur;
  }
}

// 8 values of nb bits (1..30) each -> 256-bit little-endian bit string, value k at bit k*nb (BitStuffer2.cpp:432-472)
__device__ __forceinline__ void packRow8(const uint32_t (&q)[8], int nb, uint32_t (&R)[8]) {
  const unsigned long long p0 = q[0] | ((unsigned long long)q[1] << nb), p1 = q[2] | ((unsigned long long)q[3] << nb);
  const unsigned long long p2 = q[4] | ((unsigned long long)q[5] << nb), p3 = q[6] | ((unsigned long long)q[7] << nb);
  const int s2 = 2 * nb;                                   // 2..60
  const unsigned long long h0lo = p0 | (p1 << s2), h0hi = p1 >> (64 - s2);
  const unsigned long long h1lo = p2 | (p3 << s2), h1hi = p3 >> (64 - s2);
  const int s4 = 4 * nb;                                   // 4..120
  unsigned long long r0, r1, r2, r3;
  if (s4 < 64) { r0 = h0lo | (h1lo << s4); r1 = h1lo >> (64 - s4); r2 = 0; r3 = 0; }
  else {
    const int t = s4 - 64;
    r0 = h0lo; r1 = h0hi | (h1lo << t);
    r2 = (t ? (h1lo >> (64 - t)) : 0ull) | (h1hi << t);
    r3 = t ? (h1hi >> (64 - t)) : 0ull;
  }
  R[0] = (uint32_t)r0; R[1] = (uint32_t)(r0 >> 32); R[2] = (uint32_t)r1; R[3] = (uint32_t)(r1 >> 32);
  R[4] = (uint32_t)r2; R[5] = (uint32_t)(r2 >> 32); R[6] = (uint32_t)r3; R[7] = (uint32_t)(r3 >> 32);
}

// raw bits of w values of T, value k at bit k*8*sizeof(T)
template <class T>
__device__ __forceinline__ void rawRowBits(const T (&v)[8], int w, uint32_t (&R)[16]) {
#pragma unroll
  for (int j = 0; j < 16; j++) R[j] = 0;
  if (sizeof(T) == 8) {
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < w) { unsigned long long b; memcpy(&b, &v[k], 8); R[2 * k] = (uint32_t)b; R[2 * k + 1] = (uint32_t)(b >> 32); }
  } else if (sizeof(T) == 4) {
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < w) { uint32_t b; memcpy(&b, &v[k], 4); R[k] = b; }
  } else if (sizeof(T) == 2) {
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < w) { uint16_t b; memcpy(&b, &v[k], 2); R[k >> 1] |= (uint32_t)b << (16 * (k & 1)); }
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < w) { uint8_t b; memcpy(&b, &v[k], 1); R[k >> 2] |= (uint32_t)b << (8 * (k & 3)); }
  }
}

// 8 pixels of one block row; columns >= w repeat the first pixel (neutral for min/max).
template <class T>
__device__ __forceinline__ void loadRow8(const T* __restrict__ p, int w, bool vec, T (&v)[8]) {
  if (vec) {                                          // w == 8 and 16-byte aligned
    if (sizeof(T) == 4) {
      const uint4 a = __ldg((const uint4*)p), b = __ldg((const uint4*)p + 1);
      const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 8; k++) memcpy(&v[k], &u[k], 4);
    } else if (sizeof(T) == 2) {
      const uint4 a = __ldg((const uint4*)p);
      const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int k = 0; k < 8; k++) { const uint16_t h = (uint16_t)(u[k >> 1] >> (16 * (k & 1))); memcpy(&v[k], &h, 2); }
    } else if (sizeof(T) == 8) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint4 a = __ldg((const uint4*)p + k);
        const unsigned long long lo = a.x | ((unsigned long long)a.y << 32), hi = a.z | ((unsigned long long)a.w << 32);
        memcpy(&v[2 * k], &lo, 8); memcpy(&v[2 * k + 1], &hi, 8);
      }
    } else {
      const uint2 a = __ldg((const uint2*)p);
      const uint32_t u[2] = {a.x, a.y};
#pragma unroll
      for (int k = 0; k < 8; k++) { const uint8_t h = (uint8_t)(u[k >> 2] >> (8 * (k & 3))); memcpy(&v[k], &h, 1); }
    }
  } else {
    const T first = __ldg(p);
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = k < w ? __ldg(p + k) : first;
  }
}

// ---- the kernel --------------------------------------------------------------------------------
// Persistent CTAs (256 threads = 8 warps x 4 micro-blocks).  A tile = up to 32 consecutive blocks of ONE block row
// (tiles never wrap, so a lane's pixel address needs no division); tiles are numbered in stream order and CTA c
// works on tiles c, c + gridDim.x, ...  While tile i is coded, the pixels of tile i + 1 are already in flight
// (register prefetch).  All CTAs are co-resident (the host sizes the grid by occupancy), which the look-back needs.
constexpr int FAST_TB = 32;

template <class T>
struct FastRow { T v[8]; };

// ---- general (any block shape / coding) per-lane pieces, kept out of line so that the hot path stays lean -------
template <class T>
__device__ __noinline__ void fastGenericChoice(const FastEncArgs& a, const T* __restrict__ v, int h, int w, int r, bool act,
                                               int& nBytesOut, int& modeOut, int& nbOut, int& tcOut, int& dtUsedOut, uint32_t& maxElemOut,
                                               double& zMinOut, T& loOut, unsigned int& flagsOut,
                                               typename PixelTraits<T>::Key& kminOut, typename PixelTraits<T>::Key& kmaxOut) {
  using K = typename PixelTraits<T>::Key;
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  constexpr int DT = PixelTraits<T>::code;
  const bool rowAct = act && r < h;
  K kmin = keyMaxValue<K>(), kmax = 0;
  int same = 0;
  if (rowAct) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < w) { const K key = toKey(v[k]); kmin = key < kmin ? key : kmin; kmax = key > kmax ? key : kmax; }
      if (k > 0) same += (k < w && v[k] == v[k - 1]) ? 1 : 0;
    }
  }
  {  // previous pixel of a row's first pixel: last pixel of the row above; 0 for the block's first pixel (Lerc2.cpp:1729)
    T last = v[0];
#pragma unroll
    for (int k = 1; k < 8; k++) if (k < w) last = v[k];
    T up;
    if (sizeof(T) == 8) { unsigned long long bb; memcpy(&bb, &last, 8); bb = __shfl_up_sync(FULL, bb, 1, 8); memcpy(&up, &bb, 8); }
    else { uint32_t bb = 0; memcpy(&bb, &last, sizeof(T)); bb = __shfl_up_sync(FULL, bb, 1, 8); memcpy(&up, &bb, sizeof(T)); }
    if (r == 0) up = (T)0;
    if (rowAct) same += (v[0] == up) ? 1 : 0;
  }
#pragma unroll
  for (int m = 1; m < 8; m <<= 1) {
    const K omin = shflXorK<K>(kmin, m), omax = shflXorK<K>(kmax, m);
    kmin = omin < kmin ? omin : kmin; kmax = omax > kmax ? omax : kmax;
    same += __shfl_xor_sync(FULL, same, m);
  }
  kminOut = kmin; kmaxOut = kmax;
  const int n = h * w;
  const T lo = fromKey<T>(kmin), hi = fromKey<T>(kmax);
  const double zMin = (double)lo, zMax = (double)hi;
  zMinOut = zMin; loOut = lo;
  unsigned int fl = 0;
  if (isFlt && act && (isNaNVal(lo) || isNaNVal(hi))) fl |= FASTF_NAN;
  if (act && n > 4 && (zMax > __dadd_rn(zMin, __dmul_rn(3.0, a.maxZErr))) && (2 * same > n)) fl |= FASTF_LUT;   // Lerc2.cpp:1794-1795
  flagsOut = fl;
  int mode = BEM_RAW, nb = 0, tc = 0, dtUsed = DT, nBytes = 0;
  uint32_t maxElem = 0;
  if (act) {                                                           // NumBytesTile, Lerc2.h:416-453; tryLut == false
    const int raw = 1 + n * (int)sizeof(T);
    if (zMin == 0 && zMax == 0) { nBytes = 1; mode = BEM_ZERO; }
    else {
      const double mv = __dmul_rn(__dsub_rn(zMax, zMin), a.scale);
      if (mv > (double)a.maxQ) nBytes = raw;
      else {
        tc = reduceOffsetType(zMin, DT, dtUsed);
        int nbt = 1 + dtSize(dtUsed);
        maxElem = roundToUInt(mv);
        if (maxElem > 0) { nb = bitLength(maxElem); nbt += 2 + (int)packedBytes(n, nb); }
        if (nbt < raw) { mode = BEM_SIMPLE; nBytes = nbt; } else nBytes = raw;
      }
    }
  }
  nBytesOut = nBytes; modeOut = mode; nbOut = nb; tcOut = tc; dtUsedOut = dtUsed; maxElemOut = maxElem;
}

template <class T>
__device__ __noinline__ void fastGenericEmit(const FastEncArgs& a, uint32_t* stage, const T* __restrict__ vs, int h, int w, int r, int j0, uint32_t byte0,
                                             int mode, int nb, int tc, int dtUsed, uint32_t maxElem, double zMin, T lo) {
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  if (r >= h) return;
  T v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = vs[k];
  const int n = h * w;
  const uint8_t flag = (uint8_t)((((j0 >> 3) & 15) << 2) & 0x38);       // version 6, no depth delta (Lerc2.cpp:1955-1958)
  if (mode == BEM_ZERO) { if (r == 0) { const uint32_t H[1] = {(uint32_t)(flag | 2)}; orBits<1>(stage, byte0 * 8, H, 8); } return; }
  if (mode == BEM_RAW) {
    if (r == 0) { const uint32_t H[1] = {(uint32_t)flag}; orBits<1>(stage, byte0 * 8, H, 8); }
    uint32_t R[16]; rawRowBits<T>(v, w, R);
    const uint32_t bitOff = (byte0 + 1 + (uint32_t)(r * w) * (uint32_t)sizeof(T)) * 8;
    if (sizeof(T) == 8) orBits<16>(stage, bitOff, R, w * 64);
    else { uint32_t R8[8];
#pragma unroll
           for (int j = 0; j < 8; j++) R8[j] = R[j];
           orBits<8>(stage, bitOff, R8, w * 8 * (int)sizeof(T)); }
    return;
  }
  const int osz = dtSize(dtUsed);
  if (r == 0) {                                                         // flag | offset | [numBits byte | count]
    const unsigned long long ob = offsetBits(zMin, dtUsed);
    unsigned long long lo64 = (unsigned long long)(flag | (maxElem == 0 ? 3 : 1) | (tc << 6)) | (ob << 8);
    unsigned long long hi64 = osz == 8 ? (ob >> 56) : 0;
    int nbytes = 1 + osz;                                               // 2, 3, 5 or 9
    if (maxElem > 0) {
      const unsigned long long two = (unsigned long long)(nb | (2 << 6)) | ((unsigned long long)n << 8);   // count < 256: one byte, code 2
      if (nbytes < 8) lo64 |= two << (8 * nbytes); else hi64 |= two << (8 * (nbytes - 8));
      nbytes += 2;
    }
    const uint32_t H[3] = {(uint32_t)lo64, (uint32_t)(lo64 >> 32), (uint32_t)hi64};
    orBits<3>(stage, byte0 * 8, H, nbytes * 8);
  }
  if (maxElem == 0) return;
  uint32_t q[8];
  bool done = false;
  if constexpr (!isFlt) {
    if (a.intLossless) {
#pragma unroll
      for (int k = 0; k < 8; k++) q[k] = k < w ? (uint32_t)((long long)v[k] - (long long)lo) : 0u;
      done = true;
    }
  }
  if (!done) {
#pragma unroll
    for (int k = 0; k < 8; k++) q[k] = k < w ? quantizeOne((double)v[k], zMin, a.scale) : 0u;
  }
  uint32_t R[8];
  packRow8(q, nb, R);
  orBits<8>(stage, (byte0 + (uint32_t)(osz + 3)) * 8 + (uint32_t)(r * w * nb), R, w * nb);
}

template <class T, int MINB, bool BATCH = false>
__global__ void __launch_bounds__(256, MINB) k_encode_fused(FastEncArgs a, typename std::conditional<BATCH, FastBatchArgs, FastNoBatch>::type t) {
  using K = typename PixelTraits<T>::Key;
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  constexpr int DT = PixelTraits<T>::code;
  constexpr int TB = FAST_TB;
  constexpr int MAXB = 1 + 64 * (int)sizeof(T);                 // longest block: raw (Lerc2.h:427)
  constexpr int NQ = (TB * MAXB + 15) / 16 + 3;                 // staging uint4s: 16 zero bytes | tile output | zero tail
  extern __shared__ __align__(16) uint32_t stageRaw[];          // two staging images (tiles alternate) | T sRow[256][8] (general path)
  T* sRow = (T*)(stageRaw + 2 * NQ * 4);
  __shared__ uint32_t sLen[2][TB];
  __shared__ unsigned long long sTileOff, sLocalOff;
  __shared__ unsigned long long sKMin[8], sKMax[8], sFA[8], sFD[8];
  __shared__ unsigned int sFlg[8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sb = lane >> 3, r = lane & 7;
  const unsigned int flagsSeen = BATCH ? 0u : *(volatile unsigned int*)&a.res->flags;
  const unsigned long long negMinSeen = BATCH ? 0ull : *(volatile unsigned long long*)&a.res->negMinKey, maxSeen = BATCH ? 0ull : *(volatile unsigned long long*)&a.res->maxKey;
  for (int i = tid; i < 2 * NQ; i += 256) ((uint4*)stageRaw)[i] = make_uint4(0, 0, 0, 0);

  const T* data = (const T*)a.data;
  const long long rowPitch = BATCH ? t.pitch : (long long)a.nCols;
  const bool vecOk = BATCH ? (((t.pitch * (long long)sizeof(T)) % 16 == 0) && ((t.imgCols * (int)sizeof(T)) % 16 == 0) && (((uintptr_t)data & 15) == 0))
                           : (((a.nCols * (int)sizeof(T)) % 16 == 0) && (((uintptr_t)data & 15) == 0));
  const int tpr = (a.nTx + TB - 1) / TB;                         // tiles per block row (of a full image)
  const int nTiles = BATCH ? t.segPerImg * t.nImgX * t.nImgY : tpr * a.nTy;
  const int b = warp * 4 + sb;
  constexpr unsigned long long ST_A = 1ull << 62, ST_P = 2ull << 62, VAL = (1ull << 62) - 1;
  volatile unsigned long long* st = a.tileState;

  // running image-global facts and checksum partials of this thread
  K gMin = keyMaxValue<K>(), gMax = 0;
  unsigned int myFlags = 0;
  unsigned long long fa = 0, fd = 0;
  bool overflow = false;
  uint32_t prevBytes = 0;

  // geometry + pixel row of a tile for this lane
  auto tileGeom = [&](int tile, int tyT, int segT, int& ty, int& tx, int& h, int& w, bool& act) {
    ty = tyT; tx = segT * TB + b;
    act = tile < nTiles && tx < a.nTx;
    h = act ? min(8, a.nRows - ty * 8) : 0; w = act ? min(8, a.nCols - tx * 8) : 0;
  };
  // tile batch: tile -> (image, block row, segment); `org` = element offset of the image's first pixel in the raster
  auto batchGeom = [&](int tile, int& img, int& local, size_t& org, int& ty, int& tx, int& h, int& w, bool& act) {
    img = tile / t.segPerImg; local = tile - img * t.segPerImg;
    const int tyT = local / tpr, segT = local - tyT * tpr;
    const int iy = img / t.nImgX, ix = img - iy * t.nImgX;
    const int rows = min(t.imgRows, t.rasterRows - iy * t.imgRows), cols = min(t.imgCols, t.rasterCols - ix * t.imgCols);
    ty = tyT; tx = segT * TB + b;
    act = tile < nTiles && ty * 8 < rows && tx * 8 < cols;
    h = act ? min(8, rows - ty * 8) : 0; w = act ? min(8, cols - tx * 8) : 0;
    org = (size_t)iy * t.imgRows * (size_t)t.pitch + (size_t)ix * t.imgCols;
  };
  const int stepTy = (int)gridDim.x / tpr, stepSeg = (int)gridDim.x - stepTy * tpr;   // tile += gridDim.x in (block row, segment) form
  FastRow<T> cur, nxt;
  int ty, tx, h, w; bool act;
  int tile = blockIdx.x;
  int tyT = tile / tpr, segT = tile - tyT * tpr;
  int img = 0, local = 0, nimg = 0, nlocal = 0; size_t org = 0, norg = 0;          // tile batch only
  if (BATCH) batchGeom(tile, img, local, org, ty, tx, h, w, act);
  else tileGeom(tile, tyT, segT, ty, tx, h, w, act);
#pragma unroll
  for (int k = 0; k < 8; k++) cur.v[k] = (T)0;
  if (act && r < h) loadRow8<T>(data + org + (size_t)(ty * 8 + r) * rowPitch + tx * 8, w, vecOk && w == 8, cur.v);
  __syncthreads();                                               // staging zeroed

  for (int it = 0; tile < nTiles; it++, tile += gridDim.x) {
    uint32_t* stage = stageRaw + (size_t)(it & 1) * NQ * 4 + 4;  // tile-local byte 0 of this tile's output image
    // ---- prefetch the next tile's pixels
    int nty, ntx, nh, nw; bool nact;
    if (BATCH) batchGeom(tile + gridDim.x, nimg, nlocal, norg, nty, ntx, nh, nw, nact);
    else {
      tyT += stepTy; segT += stepSeg; if (segT >= tpr) { segT -= tpr; tyT++; }
      tileGeom(tile + gridDim.x, tyT, segT, nty, ntx, nh, nw, nact);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) nxt.v[k] = (T)0;
    if (nact && r < nh) loadRow8<T>(data + norg + (size_t)(nty * 8 + r) * rowPitch + ntx * 8, nw, vecOk && nw == 8, nxt.v);

    // ---- block statistics and coding choice
    const int j0 = tx * 8, n = h * w;
    int mode = BEM_SIMPLE, nb = 0, tc = 0, dtUsed = DT, nBytes = 0; uint32_t maxElem = 0; double zMin = 0; T lo = (T)0;
    bool hot = false;
    if (isFlt && sizeof(T) == 4) {
      // hot path test for full 8x8 float blocks coded "bit-stuffed, offset as float/short/byte"
      float fv[8];
#pragma unroll
      for (int k = 0; k < 8; k++) memcpy(&fv[k], &cur.v[k], 4);
      float mn = fminf(fminf(fminf(fv[0], fv[1]), fminf(fv[2], fv[3])), fminf(fminf(fv[4], fv[5]), fminf(fv[6], fv[7])));
      float mx = fmaxf(fmaxf(fmaxf(fv[0], fv[1]), fmaxf(fv[2], fv[3])), fmaxf(fmaxf(fv[4], fv[5]), fmaxf(fv[6], fv[7])));
      float t0 = 0.f;                                              // NaN iff some value is NaN or +-Inf
#pragma unroll
      for (int k = 0; k < 8; k++) t0 = __fmaf_rn(fv[k], 0.f, t0);
      int same = 0;
#pragma unroll
      for (int k = 1; k < 8; k++) same += (fv[k] == fv[k - 1]) ? 1 : 0;
      const float up = __shfl_up_sync(FULL, fv[7], 1, 8);
      same += (fv[0] == (r == 0 ? 0.f : up)) ? 1 : 0;
      const bool full = act && h == 8 && w == 8;
      uint32_t kmn = toKey(mn), kmx = toKey(mx);
      int nonFinite = (t0 != t0) ? 1 : 0;
#pragma unroll
      for (int m = 1; m < 8; m <<= 1) {                              // 8-lane groups: xor shuffles stay inside the group
        const uint32_t omn = __shfl_xor_sync(FULL, kmn, m), omx = __shfl_xor_sync(FULL, kmx, m);
        kmn = omn < kmn ? omn : kmn; kmx = omx > kmx ? omx : kmx;
        const int pk = __shfl_xor_sync(FULL, same | (nonFinite << 16), m);
        same += pk & 0xffff; nonFinite |= pk >> 16;
      }
      const bool finite = nonFinite == 0;
      const float lof = fromKey<float>(kmn), hif = fromKey<float>(kmx);
      const double zMn = (double)lof, zMx = (double)hif;
      const double mv = __dmul_rn(__dsub_rn(zMx, zMn), a.scale);
      const uint32_t me = roundToUInt(mv);
      const int nbh = bitLength(me);
      const bool lutCand = (zMx > __dadd_rn(zMn, a.maxZErr3)) && (2 * same > 64);
      hot = full && finite && !(mv > (double)a.maxQ) && me > 0 && nbh <= 16 && !lutCand && !(lof == 0.f && hif == 0.f);
      hot = __all_sync(FULL, hot || !act) && __any_sync(FULL, act);
      if (hot && act) {
        memcpy(&lo, &lof, 4); zMin = zMn; maxElem = me; nb = nbh;
        // offset in the smallest type that holds it (Lerc2.h:457-542, float row)
        const bool isInt = lof == truncf(lof);
        tc = (isInt && lof >= 0.f && lof <= 255.f) ? 2 : ((isInt && lof >= -32768.f && lof <= 32767.f) ? 1 : 0);
        dtUsed = tc == 0 ? DT_Float : (tc == 1 ? DT_Short : DT_Byte);
        nBytes = 1 + (4 >> tc) + 2 + 8 * nb;
        gMin = kmn < gMin ? (K)kmn : gMin; gMax = kmx > gMax ? (K)kmx : gMax;
        if (!(flagsSeen & FASTF_NOT_INT) && !(myFlags & FASTF_NOT_INT)) {          // all-integer test (Lerc.h:248)
          bool ni = false;
#pragma unroll
          for (int k = 0; k < 8; k++) ni |= fv[k] != truncf(fv[k]);
          if (ni) myFlags |= FASTF_NOT_INT;
        }
      }
    }
    if (!hot) {
      unsigned int fl = 0; K kmin, kmax;
#pragma unroll
      for (int k = 0; k < 8; k++) sRow[tid * 8 + k] = cur.v[k];
      fastGenericChoice<T>(a, sRow + tid * 8, h, w, r, act, nBytes, mode, nb, tc, dtUsed, maxElem, zMin, lo, fl, kmin, kmax);
      myFlags |= fl;
      if (act) { gMin = kmin < gMin ? kmin : gMin; gMax = kmax > gMax ? kmax : gMax; }
      if (isFlt && act && r < h && !(flagsSeen & FASTF_NOT_INT) && !(myFlags & FASTF_NOT_INT)) {
        bool ni = false;
#pragma unroll
        for (int k = 0; k < 8; k++) ni |= k < w && (sizeof(T) == 4 ? ((float)cur.v[k] != truncf((float)cur.v[k])) : ((double)cur.v[k] != trunc((double)cur.v[k])));
        if (ni) myFlags |= FASTF_NOT_INT;
      }
    }
    if (r == 0) sLen[it & 1][b] = (uint32_t)nBytes;
    __syncthreads();                                             // block lengths visible (and the previous tile's flush is over)
    if (it > 0) {  // zero what the previous tile used of its image; that image is packed into again by the next tile
      const int nz = (((int)prevBytes + 15) >> 4) + 3;
      uint4* img = (uint4*)(stageRaw + (size_t)((it - 1) & 1) * NQ * 4);
      for (int i = tid; i < nz && i < NQ; i += 256) img[i] = make_uint4(0, 0, 0, 0);
    }

    // ---- tile-local byte offsets: every warp scans the 32 lengths itself
    const uint32_t myLen = sLen[it & 1][lane];
    uint32_t inc = myLen;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const uint32_t o = __shfl_up_sync(FULL, inc, s); if (lane >= s) inc += o; }
    const uint32_t tileBytes = __shfl_sync(FULL, inc, 31);
    const uint32_t byte0 = __shfl_sync(FULL, inc - myLen, b);
    if (tid == 0) st[tile] = ((BATCH ? local : tile) == 0 ? ST_P : ST_A) | (unsigned long long)tileBytes;   // publish before packing

    // ---- quantise (Lerc2.h:357-376), pack, OR into the staging image (WriteTile, Lerc2.cpp:1949-2021)
    if (hot) {
      if (act) {
        const int osz = 4 >> tc;
        if (r == 0) {                                              // flag | offset | numBits byte | count
          const uint32_t flag = (uint32_t)((((j0 >> 3) & 15) << 2) & 0x38);
          const unsigned long long ob = offsetBits(zMin, dtUsed);
          unsigned long long hd = (unsigned long long)(flag | 1 | (tc << 6)) | (ob << 8);
          hd |= ((unsigned long long)(nb | (2 << 6)) | (64ull << 8)) << (8 * (1 + osz));
          const uint32_t H[2] = {(uint32_t)hd, (uint32_t)(hd >> 32)};
          orBits<2>(stage, byte0 * 8, H, (3 + osz) * 8);
        }
        float fv[8];
#pragma unroll
        for (int k = 0; k < 8; k++) memcpy(&fv[k], &cur.v[k], 4);
        uint32_t q[8];
#pragma unroll
        for (int k = 0; k < 8; k++) q[k] = quantizeOne((double)fv[k], zMin, a.scale);
        // 8 values of nb <= 16 bits -> 128 bits
        const unsigned long long p0 = q[0] | ((unsigned long long)q[1] << nb), p1 = q[2] | ((unsigned long long)q[3] << nb);
        const unsigned long long p2 = q[4] | ((unsigned long long)q[5] << nb), p3 = q[6] | ((unsigned long long)q[7] << nb);
        const int s2 = 2 * nb, s4 = 4 * nb;
        const unsigned long long h0 = p0 | (p1 << s2), h1 = p2 | (p3 << s2);      // 4 nb <= 64 bits each
        const unsigned long long r0 = s4 == 64 ? h0 : (h0 | (h1 << s4)), r1 = s4 == 64 ? h1 : (h1 >> (64 - s4));
        const uint32_t R[4] = {(uint32_t)r0, (uint32_t)(r0 >> 32), (uint32_t)r1, (uint32_t)(r1 >> 32)};
        orBits<4>(stage, (byte0 + (uint32_t)(osz + 3) + (uint32_t)(r * nb)) * 8, R, 8 * nb);
      }
    } else if (act) {
      fastGenericEmit<T>(a, stage, sRow + tid * 8, h, w, r, j0, byte0, mode, nb, tc, dtUsed, maxElem, zMin, lo);
    }

    // ---- decoupled look-back (warp 0) for the tile's global byte offset
    if (warp == 0) {
      unsigned long long excl = 0;
      const long long first = BATCH ? (long long)(tile - local) : 0;    // the chain restarts at every image of a tile batch
      if (!BATCH && a.groupState) {
        // Two-round look-back.  With G resident CTAs the plain chain below needs ~G / 32 dependent rounds per tile (the tiles of the
        // current sweep only have aggregates yet); here round 1 covers the predecessors inside the tile's group of 32 and round 2
        // the aggregates of whole groups, published by the CTA that sizes a group's last tile.
        volatile unsigned long long* gs = a.groupState;
        const int l = tile & 31;
        const long long g = tile >> 5;
        bool needGroups = g > 0;
        {
          const long long idx = (long long)tile - 1 - lane;
          const bool in = lane < l;
          unsigned long long s = 0;
          if (in) { do { s = st[idx]; } while ((s >> 62) == 0); }
          const unsigned isP = __ballot_sync(FULL, in && (s >> 62) == 2);
          const int firstP = isP ? __ffs(isP) - 1 : 32;
          unsigned long long contrib = (in && lane <= firstP) ? (s & VAL) : 0;
#pragma unroll
          for (int m = 16; m; m >>= 1) contrib += __shfl_xor_sync(FULL, contrib, m);
          excl = contrib;
          if (isP) needGroups = false;                                  // an inclusive prefix inside the group: excl is already global
        }
        if (l == 31 && needGroups && lane == 0) gs[g] = ST_A | (excl + tileBytes);    // this group's bytes (its 32 tiles are all sized)
        if (needGroups) {
          long long base = g - 1;
          for (;;) {
            const long long idx = base - lane;
            unsigned long long s = ST_P;                                // virtual groups before 0: prefix 0
            if (idx >= 0) { do { s = gs[idx]; } while ((s >> 62) == 0); }
            const unsigned isP = __ballot_sync(FULL, (s >> 62) == 2);
            const int firstP = isP ? __ffs(isP) - 1 : 32;
            unsigned long long contrib = (lane <= firstP && idx >= 0) ? (s & VAL) : 0;
#pragma unroll
            for (int m = 16; m; m >>= 1) contrib += __shfl_xor_sync(FULL, contrib, m);
            excl += contrib;
            if (isP) break;
            base -= 32;
          }
        }
        if (lane == 0) {
          if (tile > 0) st[tile] = ST_P | (excl + tileBytes);
          if (l == 31) gs[g] = ST_P | (excl + tileBytes);               // inclusive prefix of the whole group
        }
      } else
      if (tile > first) {
        long long base = (long long)tile - 1;
        for (;;) {
          const long long idx = base - lane;
          unsigned long long s = ST_P;                                  // virtual tiles before the first: prefix 0
          if (idx >= first) { do { s = st[idx]; } while ((s >> 62) == 0); }
          const unsigned isP = __ballot_sync(FULL, (s >> 62) == 2);
          const int firstP = isP ? __ffs(isP) - 1 : 32;
          unsigned long long contrib = (lane <= firstP && idx >= first) ? (s & VAL) : 0;
#pragma unroll
          for (int m = 16; m; m >>= 1) contrib += __shfl_xor_sync(FULL, contrib, m);
          excl += contrib;
          if (isP) break;
          base -= 32;
        }
        if (lane == 0) st[tile] = ST_P | (excl + tileBytes);
      }
      if (!BATCH) {
        if (lane == 0) {
          sTileOff = excl;
          if (tile == nTiles - 1) a.res->totalBytes = excl + tileBytes;
        }
      } else {
        // second look-back, over whole blobs: where does this image's blob start in the output?
        volatile unsigned long long* ist = t.imgState;
        const bool lastSeg = local == t.segPerImg - 1;
        const unsigned long long blobBytes = (unsigned long long)t.dataStart + excl + tileBytes;      // meaningful for the last segment only
        if (lastSeg && lane == 0) { t.imgRes[img].streamBytes = excl + tileBytes; __threadfence(); ist[img] = (img == 0 ? ST_P : ST_A) | blobBytes; }
        unsigned long long imgBase = 0;
        if (img > 0) {
          long long base = (long long)img - 1;
          for (;;) {
            const long long idx = base - lane;
            unsigned long long s = ST_P;
            if (idx >= 0) { do { s = ist[idx]; } while ((s >> 62) == 0); }
            const unsigned isP = __ballot_sync(FULL, (s >> 62) == 2);
            const int firstP = isP ? __ffs(isP) - 1 : 32;
            unsigned long long contrib = (lane <= firstP && idx >= 0) ? (s & VAL) : 0;
#pragma unroll
            for (int m = 16; m; m >>= 1) contrib += __shfl_xor_sync(FULL, contrib, m);
            imgBase += contrib;
            if (isP) break;
            base -= 32;
          }
          if (lastSeg && lane == 0) ist[img] = ST_P | (imgBase + blobBytes);
        }
        if (lane == 0) { sTileOff = imgBase + (unsigned long long)t.dataStart + excl; sLocalOff = excl; }
      }
    }
    __syncthreads();                                             // staging image complete, tile offset known

    // ---- staging -> HBM in 16-byte chunks aligned to the GLOBAL address (the staging image is re-aligned with
    // funnel shifts), Fletcher-32 partial sums of the same words (bytes outside the tile are zero in the image);
    // every chunk read is zeroed again for the tile after next
    const unsigned long long tileOff = sTileOff;
    const unsigned long long localOff = BATCH ? sLocalOff : tileOff;     // offset inside this blob's micro-block stream
    {
      uint8_t* gTile = (BATCH ? t.out : a.stream) + tileOff;
      const bool fits = tileOff + tileBytes <= (BATCH ? t.outCap : a.streamCap);
      if (!fits) overflow = true;
      const int pad = (int)((uintptr_t)gTile & 15);
      const int nChunks = (pad + (int)tileBytes + 15) >> 4;
      const int bs8 = ((-pad) & 3) * 8;
      for (int cI = tid; cI < nChunks; cI += 256) {
        const int s0 = cI * 16 - pad;                                     // tile-local byte of the chunk's first byte (>= -15)
        const int wi = s0 >> 2;                                           // floor
        uint32_t x[5];
#pragma unroll
        for (int k = 0; k < 5; k++) x[k] = stage[wi + k];
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = __funnelshift_r(x[k], x[k + 1], bs8);
        if (fits) {
          if (s0 >= 0 && s0 + 16 <= (int)tileBytes) *(uint4*)(gTile + s0) = make_uint4(o[0], o[1], o[2], o[3]);
          else {
#pragma unroll
            for (int j = 0; j < 16; j++) if (s0 + j >= 0 && s0 + j < (int)tileBytes) gTile[s0 + j] = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
          }
        }
        // big-endian 16-bit words at even region offsets (Lerc2.cpp:1037-1064)
        const long long r0 = a.regionOff + (long long)localOff + s0;
        const unsigned par = (unsigned)(r0 & 1);
        const uint32_t w0 = (uint32_t)((unsigned long long)(r0 + 16) >> 1) % 65535u + 65535u - 8u;   // word index of byte r0 - par (mod 65535)
        uint32_t S = 0, S1 = 0, prev = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
          const uint32_t cw = k < 4 ? o[k] : 0u;
          const uint32_t y = __funnelshift_l(prev, cw, par * 8);        // bytes shifted up by one when the chunk starts at an odd offset
          const uint32_t pw = __byte_perm(y, 0, 0x2301);                // low half = first BE word, high half = second
          const uint32_t wlo = pw & 0xffffu, whi = pw >> 16;
          S += wlo + whi; S1 += (uint32_t)(2 * k) * (wlo + whi) + whi;
          prev = cw;
        }
        fa += S; fd += (unsigned long long)w0 * S + S1;
      }
    }
    prevBytes = tileBytes;
    if (BATCH && (nimg != img || tile + (int)gridDim.x >= nTiles)) {
      // ---- the next tile belongs to another image: hand this image's facts and checksum partials over
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) {
        const K omin = shflXorK<K>(gMin, m), omax = shflXorK<K>(gMax, m);
        gMin = omin < gMin ? omin : gMin; gMax = omax > gMax ? omax : gMax;
      }
      if (overflow) myFlags |= FASTF_OVERFLOW;
      myFlags = __reduce_or_sync(FULL, myFlags);
      fa %= 65535ull; fd %= 65535ull;
#pragma unroll
      for (int m = 16; m; m >>= 1) { fa += __shfl_xor_sync(FULL, fa, m); fd += __shfl_xor_sync(FULL, fd, m); }
      if (lane == 0) { sKMin[warp] = (unsigned long long)gMin; sKMax[warp] = (unsigned long long)gMax; sFlg[warp] = myFlags; sFA[warp] = fa; sFD[warp] = fd; }
      __syncthreads();
      if (tid == 0) {
        unsigned long long A = 0, D = 0, kMin = ~0ull, kMax = 0; unsigned int fl = 0;
        for (int i = 0; i < 8; i++) { A += sFA[i]; D += sFD[i]; kMin = sKMin[i] < kMin ? sKMin[i] : kMin; kMax = sKMax[i] > kMax ? sKMax[i] : kMax; fl |= sFlg[i]; }
        TileEncResult* ir = t.imgRes + img;
        if (A | D) { atomicAdd(&ir->fletA, A); atomicAdd(&ir->fletD, D % 65535ull); }
        if (kMax >= kMin) { atomicMax(&ir->negMinKey, ~kMin); atomicMax(&ir->maxKey, kMax); }
        if (fl) atomicOr(&ir->flags, fl);
      }
      gMin = keyMaxValue<K>(); gMax = 0; myFlags = 0; fa = 0; fd = 0; overflow = false;
      // (the barrier after the next tile's block lengths separates these shared arrays from their next use)
    }
    // ---- next tile
    cur = nxt; ty = nty; tx = ntx; h = nh; w = nw; act = nact;
    if (BATCH) { img = nimg; local = nlocal; org = norg; }
  }
  if (BATCH) return;

  // ---- image-global facts and checksum partials of this CTA
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    const K omin = shflXorK<K>(gMin, m), omax = shflXorK<K>(gMax, m);
    gMin = omin < gMin ? omin : gMin; gMax = omax > gMax ? omax : gMax;
  }
  if (overflow) myFlags |= FASTF_OVERFLOW;
  myFlags = __reduce_or_sync(FULL, myFlags);
  fa %= 65535ull; fd %= 65535ull;
#pragma unroll
  for (int m = 16; m; m >>= 1) { fa += __shfl_xor_sync(FULL, fa, m); fd += __shfl_xor_sync(FULL, fd, m); }
  if (lane == 0) { sKMin[warp] = (unsigned long long)gMin; sKMax[warp] = (unsigned long long)gMax; sFlg[warp] = myFlags; sFA[warp] = fa; sFD[warp] = fd; }
  __syncthreads();
  if (tid == 0) {
    unsigned long long A = 0, D = 0, kMin = ~0ull, kMax = 0; unsigned int fl = 0;
    for (int i = 0; i < 8; i++) { A += sFA[i]; D += sFD[i]; kMin = sKMin[i] < kMin ? sKMin[i] : kMin; kMax = sKMax[i] > kMax ? sKMax[i] : kMax; fl |= sFlg[i]; }
    if (A | D) { atomicAdd(&a.res->fletA[blockIdx.x % FAST_SLOTS], A); atomicAdd(&a.res->fletD[blockIdx.x % FAST_SLOTS], D % 65535ull); }
    if (kMax >= kMin) {
      if (~kMin > negMinSeen) atomicMax(&a.res->negMinKey, ~kMin);
      if (kMax > maxSeen) atomicMax(&a.res->maxKey, kMax);
    }
    if (fl & ~flagsSeen) atomicOr(&a.res->flags, fl);
  }
}

// small blob prefix (header, mask length, ranges, flag bytes) written from kernel parameters
struct PrefixBytes { uint8_t b[128]; int n; };
__global__ void k_write_prefix(uint8_t* dst, PrefixBytes p) { if ((int)threadIdx.x < p.n) dst[threadIdx.x] = p.b[threadIdx.x]; }

}  // namespace lerc
