// lerc_encode_tile.cuh -- the single-pass Lerc2 band encoder of the headline path (included by lerc_encode.cu).
// Raster shape: every pixel valid, nDepth == 1, 8x8 micro-blocks, 16/32/64-bit pixel types.
//
// Persistent, warp-specialised CTAs (8 compute warps + 1 control warp, several CTAs per SM).  A CTA codes one tile after the
// other; a tile = 8 pixel rows x TW micro-blocks of one block row, taken by an atomic ticket in stream order
// (Lerc2.cpp:1507-1521), so forward progress never depends on which CTAs are co-resident.  Per tile k:
//
//   stage     the tile's rows are copied into shared memory by the copy engine: one cp.async.bulk (TMA, SASS UBLKCP) per
//             pixel row, completion on an mbarrier (lerc_tma.cuh).  The copy of tile k+1 (whose ticket was requested at the
//             top of tile k) is issued as soon as tile k is packed and runs under the flush.  The row pitch is the row size
//             + 16 bytes, which makes both access patterns below bank-conflict free.
//   size      two threads per micro-block: min / max / non-finite / equal-neighbour filter over its 64 pixels
//             (GetValidDataAndStats, Lerc2.cpp:1717-1799), coding choice + byte length (NumBytesTile, Lerc2.h:416-453)
//   scan      warp 0: exclusive scan of the block lengths; the tile's byte count is published for the other CTAs and
//             posted to the control warp
//   look-back control warp: decoupled look-back over the tiles' byte counts (two levels: predecessors inside the group of
//             32 tiles, then aggregates of whole groups) gives the tile's byte offset in the stream.  Nobody waits for
//             it: the offset of tile k is needed when tile k+1 has been packed.
//   pack      one thread per block row: fp64 quantisation without contraction (Quantize, Lerc2.h:357-376), 8 x numBits
//             bits packed in registers, OR-ed into one of two staging images of the tile's output bytes
//             (WriteTile Lerc2.cpp:1949-2021, BitStuffer2::EncodeSimple BitStuffer2.cpp:35-75, :432-472)
//   flush     of tile k-1: staging -> HBM in 16-byte chunks aligned to the GLOBAL address (the image is re-aligned with
//             funnel shifts), Fletcher-32 partial sums of exactly those bytes with dp4a (Lerc2.cpp:1037-1064)
//
// Like its predecessor the kernel is speculative about the image-global decisions (Lerc2.cpp:179-381); the facts it
// collects (min / max, NaN, all-integer, LUT candidates, overflow) let the caller verify them.
#pragma once
#include "lerc_tma.cuh"
#include "lerc_lookback.cuh"
#include "lerc_fletcher.cuh"

namespace lerc {

template <class T> struct EncTile {
  static constexpr int ROWB = 8 * (int)sizeof(T);                 // bytes of one block row
  static constexpr int TW = (4096 / ROWB) < 128 ? (4096 / ROWB) : 128;   // micro-blocks per tile: 128 (16/32-bit), 64 (64-bit)
  static constexpr int ROW_BYTES = TW * ROWB;                     // 4096 or 2048
  static constexpr int PITCH = ROW_BYTES + 16;                    // == 16 mod 128: rows r = 0..7 of a block fall into 8 different 16-byte bank groups
  static constexpr int IN_BYTES = 8 * PITCH;
  static constexpr int STAGE_CAP = 17920;                         // output bytes coded per pass (128 blocks of 16-bit values: 17280)
  static constexpr int STAGE_BYTES = 16 + STAGE_CAP + 48;         // 16 zero bytes | image | zero tail
  static constexpr int INFO_OFF = IN_BYTES + 2 * STAGE_BYTES;     // two staging images (tiles alternate), then uint4 sInfo[TW]
  static constexpr int OFFS_OFF = INFO_OFF + TW * 16;             // uint32 sOff[TW + 1]
  static constexpr int SMEM = OFFS_OFF + (TW + 1) * 4 + 12;
};

// sInfo[b]: x = low 32 bits of the block minimum, w = high 32 bits (64-bit types), z = unused,
// y = numBits | offset type code << 5 | mode << 7 | (maxElem == 0) << 9 | hot << 10
enum { TINFO_HOT = 1 << 10, TINFO_CONST = 1 << 9 };

// ---- coding choice of one block by ONE thread, any block shape / pixel type (the path of everything the hot path excludes:
// edge blocks, raw / all-zero / constant blocks, LUT candidates, values wider than 16 bits, non-float types)
template <class T>
__device__ __noinline__ void tileGenericChoice(const FastEncArgs& a, const uint8_t* __restrict__ blk, int pitch, int h, int w,
                                               uint4& info, uint32_t& len, unsigned int& flags,
                                               typename PixelTraits<T>::Key& kminOut, typename PixelTraits<T>::Key& kmaxOut) {
  using K = typename PixelTraits<T>::Key;
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  constexpr int DT = PixelTraits<T>::code;
  K kmin = keyMaxValue<K>(), kmax = 0;
  int same = 0;
  bool notInt = false;
  T prev = (T)0;                                                        // Lerc2.cpp:1729
  for (int y = 0; y < h; y++) {
    const T* row = (const T*)(blk + (size_t)y * pitch);
    for (int x = 0; x < w; x++) {
      const T v = row[x];
      const K key = toKey(v);
      kmin = key < kmin ? key : kmin; kmax = key > kmax ? key : kmax;
      same += (v == prev) ? 1 : 0; prev = v;
      if (isFlt) notInt |= (sizeof(T) == 4 ? ((float)v != truncf((float)v)) : ((double)v != trunc((double)v)));
    }
  }
  kminOut = kmin; kmaxOut = kmax;
  const int n = h * w;
  const T lo = fromKey<T>(kmin), hi = fromKey<T>(kmax);
  const double zMin = (double)lo, zMax = (double)hi;
  unsigned int fl = notInt ? FASTF_NOT_INT : 0;
  if (isFlt && (isNaNVal(lo) || isNaNVal(hi))) fl |= FASTF_NAN;
  if (n > 4 && (zMax > __dadd_rn(zMin, a.maxZErr3)) && (2 * same > n)) fl |= FASTF_LUT;            // Lerc2.cpp:1794-1795
  flags = fl;
  int mode = BEM_RAW, nb = 0, tc = 0, dtUsed = DT, nBytes = 0;
  uint32_t maxElem = 0;
  const int raw = 1 + n * (int)sizeof(T);                                                          // NumBytesTile, Lerc2.h:416-453; tryLut == false
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
  unsigned long long lb = 0; memcpy(&lb, &lo, sizeof(T));
  info.x = (uint32_t)lb; info.w = (uint32_t)(lb >> 32); info.z = 0;
  info.y = (uint32_t)(nb | (tc << 5) | (mode << 7) | (maxElem == 0 ? TINFO_CONST : 0));
  len = (uint32_t)nBytes;
}

// OR a word into shared memory: one ATOMS.OR without a result.  (ptxas turns a predicated red into a branch around the ATOMS,
// four instructions; OR-ing a zero word costs one.)
__device__ __forceinline__ void smemOr(uint32_t* p, uint32_t v) {
#ifndef LERC_CUSIM
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(smemAddr(p)), "r"(v) : "memory");
#else
  atomicOr(p, v);
#endif
}

// Row 0 of the image against the coarser decimal grids of Lerc2::TryRaiseMaxZError (Lerc2.cpp:1233-1339): per candidate factor the
// largest |round(x * fac) - x * fac| of the tile's part of row 0 (the tiles of block row 0 call this; the finish compares).
template <class T>
__device__ __noinline__ void encRaiseRow0(const FastEncArgs& a, const T* __restrict__ row, int cols, int tid, int nThreads) {
  double m[9];
#pragma unroll
  for (int c = 0; c < 9; c++) m[c] = 0;
  for (int e = tid; e < cols; e += nThreads) {
    const double x = (double)row[e];
#pragma unroll
    for (int c = 0; c < 9; c++) {
      if (c < a.nRaise) {
        const double z = __dmul_rn(x, a.raiseFac[c]);
        const double dlt = fabs(__dsub_rn(floor(__dadd_rn(z, 0.5)), z));
        if (dlt > m[c]) m[c] = dlt;           // NaN / Inf never win, as with std::max(a, NaN)
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 9; c++) {
    if (c < a.nRaise) {
      unsigned long long b = (unsigned long long)__double_as_longlong(m[c]);     // non-negative doubles order like integers
      for (int s = 16; s; s >>= 1) { const unsigned long long o = __shfl_xor_sync(FULL, b, s); b = o > b ? o : b; }
      if ((tid & 31) == 0 && b) atomicMax(&a.res->raiseMax[c], b);
    }
  }
}

constexpr int ENC_COMPUTE = 256, ENC_THREADS = ENC_COMPUTE + 32;    // 8 compute warps + the control warp

// BATCH (tile batch, lerc_tiles_encode.cuh): the raster is cut into equal images, every image its own blob.  A tile is then TW
// consecutive micro-blocks of ONE image in stream order = TW / nTx whole block rows (the host guarantees nTx | TW), staged side by
// side in shared memory so that sizing and packing see the same 8-row strip as for a single image.  All tiles of all images
// form one look-back chain; a blob starts at (bytes of all earlier tiles) + img * dataStart.  Per-image facts and checksum partials
// go to fb.imgRes by atomics, tile by tile.
template <class T, int MINB, bool BATCH = false>
__global__ void __launch_bounds__(ENC_THREADS, MINB) k_encode_tile(FastEncArgs a, FastBatchArgs fb) {
  using K = typename PixelTraits<T>::Key;
  using C = EncTile<T>;
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  constexpr bool hotType = isFlt && sizeof(T) == 4;
  constexpr int TW = C::TW, PITCH = C::PITCH, ROWB = C::ROWB;
  extern __shared__ __align__(16) uint8_t tileSmem[];
  uint8_t* sIn = tileSmem;
  uint4* sInfo = (uint4*)(tileSmem + C::INFO_OFF);
  uint32_t* sOff = (uint32_t*)(tileSmem + C::OFFS_OFF);
  __shared__ __align__(8) uint64_t sBarFull, sBarScan[2], sBarOff[2];
  __shared__ int sTileNext, sMailTile[2];
  __shared__ unsigned long long sMailBytes[2], sOffS[2];
  __shared__ long long sRegS[2];                                  // checksum-region offset of the tile's first output byte
  __shared__ unsigned long long sKMin[8], sKMax[8], sFA[8], sFD[8];
  __shared__ unsigned int sFlg[8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* data = (const T*)a.data;
  const bool vecOk = BATCH || ((((long long)a.nCols * (int)sizeof(T)) % 16 == 0) && (((uintptr_t)data & 15) == 0));
  const int tpr = (a.nTx + TW - 1) / TW;                          // tiles per block row
  const int nTiles = BATCH ? a.tileEnd : tpr * a.nTy;
  const int rowsPerTile = BATCH ? TW / a.nTx : 1;                 // block rows of an image in one tile (BATCH)
  volatile unsigned long long* st = a.tileState;

  // the copy engine brings tile t into sIn (thread 0 of the compute warps)
  auto issueLoad = [&](int t) {
    if (BATCH) {
      const int img = t / fb.tilesPerImg, ti = t - img * fb.tilesPerImg;
      const int iy = img / fb.nImgX, ix = img - iy * fb.nImgX;
      const int ry0 = ti * rowsPerTile, rows = min(rowsPerTile, a.nTy - ry0);
      const uint32_t rowBytes = (uint32_t)a.nTx * (uint32_t)ROWB;
      mbarExpectTx(&sBarFull, rowBytes * 8u * (uint32_t)rows);
      const T* src = data + ((size_t)iy * fb.imgRows + (size_t)ry0 * 8) * (size_t)fb.pitch + (size_t)ix * fb.imgCols;
      for (int ry = 0; ry < rows; ry++)
        for (int y = 0; y < 8; y++) bulkLoad(sIn + y * PITCH + ry * (int)rowBytes, src + (size_t)(ry * 8 + y) * (size_t)fb.pitch, rowBytes, &sBarFull);
      mbarSimCopiesDone(&sBarFull);
      return;
    }
    const int tyT = t / tpr, seg = t - tyT * tpr;
    const int h = min(8, a.nRows - tyT * 8), cols = min(TW * 8, a.nCols - seg * TW * 8);
    const uint32_t rowBytes = (uint32_t)cols * (uint32_t)sizeof(T);
    mbarExpectTx(&sBarFull, rowBytes * (uint32_t)h);
    const T* src = data + (size_t)(tyT * 8) * a.nCols + (size_t)seg * TW * 8;
    for (int y = 0; y < h; y++) bulkLoad(sIn + y * PITCH, src + (size_t)y * a.nCols, rowBytes, &sBarFull);
    mbarSimCopiesDone(&sBarFull);
  };

  if (tid == 0) {
    mbarInit(&sBarFull, 1); mbarInit(&sBarScan[0], 1); mbarInit(&sBarScan[1], 1); mbarInit(&sBarOff[0], 1); mbarInit(&sBarOff[1], 1);
    const int t = a.tileBegin + (int)atomicAdd(a.ticket, 1u);
    sTileNext = t;
    if (vecOk && t < a.tileEnd) issueLoad(t);
  }
  for (int i = tid; i < 2 * C::STAGE_BYTES / 16; i += ENC_THREADS) ((uint4*)(tileSmem + C::IN_BYTES))[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  // ================= control warp: look-back of every tile the compute warps post =================
  if (warp == ENC_COMPUTE / 32) {
    volatile unsigned long long* gs = a.groupState;
    for (int k = 0;; k++) {
      mbarWait(&sBarScan[k & 1], (uint32_t)(k >> 1) & 1u);
      const int tile = sMailTile[k & 1];
      if (tile < 0) break;
      const unsigned long long tileBytes = sMailBytes[k & 1];
      const unsigned long long excl = lookbackExclusive(st, a.groupAcc, gs, tile, tileBytes, lane);
      if (BATCH) {
        // where the tile's bytes go and where they sit inside their image's blob: the image's earlier tiles have all published their lengths
        __threadfence();
        const int img = tile / fb.tilesPerImg, first = img * fb.tilesPerImg;
        uint32_t inImg = 0;
        for (int j = first + lane; j < tile; j += 32) inImg += *(volatile const uint32_t*)&fb.tileLen[j];
        inImg = __reduce_add_sync(FULL, inImg);
        if (lane == 0) {
          sOffS[k & 1] = excl + (unsigned long long)(img + 1) * (unsigned long long)fb.dataStart;
          sRegS[k & 1] = (long long)fb.dataStart - 14 + (long long)inImg;
          mbarArrive(&sBarOff[k & 1]);
        }
        __syncwarp();
        continue;
      }
      if (lane == 0) {
        if (tile == a.tileEnd - 1 && a.hostEnd) *(volatile unsigned long long*)a.hostEnd = excl + tileBytes;      // (mapped host memory: read by the host after this launch's event)
        if (tile == nTiles - 1) { a.res->totalBytes = excl + tileBytes; __threadfence(); *(volatile unsigned int*)&a.res->totalReady = 1u; }
        sOffS[k & 1] = excl; sRegS[k & 1] = a.regionOff + (long long)excl;
        mbarArrive(&sBarOff[k & 1]);
      }
      __syncwarp();
    }
    return;
  }

  // ================= compute warps =================
  const unsigned int flagsSeen = BATCH ? 0u : *(volatile unsigned int*)&a.res->flags;
  const unsigned long long negMinSeen = *(volatile unsigned long long*)&a.res->negMinKey, maxSeen = *(volatile unsigned long long*)&a.res->maxKey;
  // running image-global facts and checksum partials of this thread
  K gMin = keyMaxValue<K>(), gMax = 0;
  unsigned int myFlags = 0;
  unsigned long long fa = 0, fd = 0;
  bool overflow = false;

  // staging image -> HBM: passBytes bytes that start at stream offset passOff
  // (regionStart: checksum-region offset of the pass's first byte; img: the image the bytes belong to, BATCH)
  auto flushPass = [&](const uint32_t* stage, unsigned long long passOff, long long regionStart, uint32_t passBytes, int img) {
    uint8_t* gPass = a.stream + passOff;
    const bool fits = passOff + passBytes <= a.streamCap;
    if (!fits) overflow = true;
    const int pad = (int)((uintptr_t)gPass & 15);
    const unsigned par = (unsigned)((regionStart - pad) & 1);         // parity of the region offset of the 16-byte aligned output bytes
    const int nChunks = (pad + (int)passBytes + 15) >> 4;
    const int bs8 = ((-pad) & 3) * 8;
    for (int cI = tid; cI < nChunks; cI += ENC_COMPUTE) {
      const int s0 = cI * 16 - pad;                                     // pass-local byte of the chunk's first byte (>= -15)
      const int wi = s0 >> 2;                                           // floor
      uint32_t x[5];
#pragma unroll
      for (int k = 0; k < 5; k++) x[k] = stage[wi + k];
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; k++) o[k] = __funnelshift_r(x[k], x[k + 1], bs8);
      if (fits) {
        if (s0 >= 0 && s0 + 16 <= (int)passBytes) *(uint4*)(gPass + s0) = make_uint4(o[0], o[1], o[2], o[3]);
        else {
#pragma unroll
          for (int j = 0; j < 16; j++) if (s0 + j >= 0 && s0 + j < (int)passBytes) gPass[s0 + j] = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
        }
      }
      const long long r0 = regionStart + s0;                            // region offset of the chunk's first byte; r0 & 1 == par
      const uint32_t w0 = (uint32_t)((unsigned long long)(r0 - par) >> 1) % 65535u;   // word index of the chunk's first word (mod 65535)
      uint32_t S, S1;
      if (par) fletcherChunk<1>(o, S, S1); else fletcherChunk<0>(o, S, S1);
      fa += S; fd += (unsigned long long)w0 * S + S1;
    }
    if (BATCH) {                                                         // the image's checksum partials and overflow flag
      uint32_t A32 = (uint32_t)fa, D32 = (uint32_t)(fd % 65535ull);      // (a thread sums at most a few chunks per pass: fa < 2^24)
      A32 = __reduce_add_sync(FULL, A32); D32 = __reduce_add_sync(FULL, D32);
      if (lane == 0) {
        TileEncResult* r = &fb.imgRes[img];
        if (A32 | D32) { atomicAdd(&r->fletA, (unsigned long long)A32); atomicAdd(&r->fletD, (unsigned long long)D32); }
        if (!fits) atomicOr(&r->flags, (unsigned int)FASTF_OVERFLOW);
      }
      fa = 0; fd = 0;
    }
  };
  // headers + rows of blocks [bLo, bHi) into `stage` (pass-local byte 0 = tile byte passBase)
  auto packBlocks = [&](uint32_t* stage, int bx0, int h, int bLo, int bHi, uint32_t passBase) {
    // headers of the hot blocks: flag | offset | numBits byte | count (WriteTile, Lerc2.cpp:1949-2021; BitStuffer2.cpp:35-75)
    for (int b = bLo + tid; b < bHi; b += ENC_COMPUTE) {
      const uint4 info = sInfo[b];
      if (info.y & TINFO_HOT) {
        const int nb = info.y & 31, tc = (info.y >> 5) & 3, osz = 4 >> tc;
        const int j0 = BATCH ? (b & (a.nTx - 1)) * 8 : (bx0 + b) * 8;
        const uint32_t flag = (uint32_t)((((j0 >> 3) & 15) << 2) & 0x38);
        const float lof = __uint_as_float(info.x);
        const unsigned long long ob = tc == 0 ? (unsigned long long)info.x : (tc == 1 ? (unsigned long long)(uint16_t)(int16_t)lof : (unsigned long long)(uint8_t)lof);
        unsigned long long hd = (unsigned long long)(flag | 1 | (tc << 6)) | (ob << 8);
        hd |= ((unsigned long long)(nb | (2 << 6)) | (64ull << 8)) << (8 * (1 + osz));
        const uint32_t H[2] = {(uint32_t)hd, (uint32_t)(hd >> 32)};
        orBits<2>(stage, (sOff[b] - passBase) * 8, H, (3 + osz) * 8);
      }
    }
    // block rows: 8 lanes per block
    for (int g = bLo * 8 + tid; g < bHi * 8; g += ENC_COMPUTE) {
      const int bb = g >> 3, r = g & 7;
      const uint4 info = sInfo[bb];
      const uint32_t byte0 = sOff[bb] - passBase;
      const uint8_t* row = sIn + r * PITCH + bb * ROWB;
      if (hotType && (info.y & TINFO_HOT)) {
        const int nb = info.y & 31, osz = 4 >> ((info.y >> 5) & 3);
        const uint4 A = *(const uint4*)row, B = *(const uint4*)(row + 16);
        const double zMin = (double)__uint_as_float(info.x);
        uint32_t q[8];
        q[0] = quantizeOne((double)__uint_as_float(A.x), zMin, a.scale); q[1] = quantizeOne((double)__uint_as_float(A.y), zMin, a.scale);
        q[2] = quantizeOne((double)__uint_as_float(A.z), zMin, a.scale); q[3] = quantizeOne((double)__uint_as_float(A.w), zMin, a.scale);
        q[4] = quantizeOne((double)__uint_as_float(B.x), zMin, a.scale); q[5] = quantizeOne((double)__uint_as_float(B.y), zMin, a.scale);
        q[6] = quantizeOne((double)__uint_as_float(B.z), zMin, a.scale); q[7] = quantizeOne((double)__uint_as_float(B.w), zMin, a.scale);
        // 8 values of nb <= 16 bits -> 128 bits, value k at bit k * nb (BitStuffer2.cpp:432-472): pairs by multiply-add, then clamped funnel shifts
        const uint32_t m = 1u << nb;
        const uint32_t p0 = q[1] * m + q[0], p1 = q[3] * m + q[2], p2 = q[5] * m + q[4], p3 = q[7] * m + q[6];   // 2 nb <= 32 bits each
        const int s2 = 2 * nb;
        const uint32_t a0 = p0 | __funnelshift_lc(0u, p1, s2), a1 = __funnelshift_lc(p1, 0u, s2);        // values 0..3: 4 nb <= 64 bits
        const uint32_t b0 = p2 | __funnelshift_lc(0u, p3, s2), b1 = __funnelshift_lc(p3, 0u, s2);        // values 4..7
        uint32_t R0, R1, R2, R3;
        if (nb > 8) { const int t = 4 * nb - 32; R0 = a0; R1 = a1 | __funnelshift_lc(0u, b0, t); R2 = __funnelshift_lc(b0, b1, t); R3 = __funnelshift_lc(b1, 0u, t); }
        else { const int t = 4 * nb; R0 = a0 | __funnelshift_lc(0u, b0, t); R1 = __funnelshift_lc(b0, 0u, t); R2 = 0; R3 = 0; }
        // the row's nb bytes start at a byte position: shift into place, OR the words into the image
        const uint32_t dByte = byte0 + (uint32_t)(osz + 3) + (uint32_t)(r * nb);
        const uint32_t sh = (dByte & 3) * 8;
        uint32_t* wp = stage + (dByte >> 2);
        smemOr(wp, R0 << sh);
        smemOr(wp + 1, __funnelshift_l(R0, R1, sh));
        smemOr(wp + 2, __funnelshift_l(R1, R2, sh));
        smemOr(wp + 3, __funnelshift_l(R2, R3, sh));                // (zero words included: the image has a zero tail behind the pass's bytes)
        const uint32_t x4 = __funnelshift_l(R3, 0u, sh);            // non-zero only for 15- and 16-bit values
        if (x4) smemOr(wp + 4, x4);
      } else {
        const int xb = BATCH ? (bb & (a.nTx - 1)) : bx0 + bb;      // block column inside the image
        const int w = min(8, a.nCols - xb * 8);
        const int mode = (info.y >> 7) & 3, nb = info.y & 31, tc = (info.y >> 5) & 3;
        const int dtUsed = offsetTypeFromCode(PixelTraits<T>::code, tc);
        const unsigned long long lb = (unsigned long long)info.x | ((unsigned long long)info.w << 32);
        T lo; memcpy(&lo, &lb, sizeof(T));
        fastGenericEmit<T>(a, stage, (const T*)row, h, w, r, xb * 8, byte0, mode, nb, tc, dtUsed, (info.y & TINFO_CONST) ? 0u : 1u, (double)lo, lo);
      }
    }
  };

  int pendK = -1, pendImg = 0; uint32_t pendBytes = 0;             // the tile whose staging image still waits for its offset
  int tile = sTileNext;
  int k = 0;                                                        // tiles this CTA has posted to its control warp
  for (; tile < a.tileEnd; k++) {
    int nextT = 0;
    const int img = BATCH ? tile / fb.tilesPerImg : 0;
    const int tyT = BATCH ? 1 : tile / tpr, seg = BATCH ? 0 : tile - tyT * tpr;
    const int bx0 = seg * TW;                                      // first block column of the tile (BATCH: block b of the tile is column b mod nTx)
    const int nbk = BATCH ? min(rowsPerTile, a.nTy - (tile - img * fb.tilesPerImg) * rowsPerTile) * a.nTx : min(TW, a.nTx - bx0);   // blocks in the tile
    const int h = BATCH ? 8 : min(8, a.nRows - tyT * 8);
    uint32_t* stage = (uint32_t*)(tileSmem + C::IN_BYTES + (k & 1) * C::STAGE_BYTES) + 4;   // pass-local byte 0 of this tile's output image
    if (vecOk) mbarWait(&sBarFull, (uint32_t)k & 1u);
    else {
      const int cols = min(TW * 8, a.nCols - bx0 * 8);
      const T* src = data + (size_t)(tyT * 8) * a.nCols + (size_t)bx0 * 8;
      for (int y = 0; y < h; y++)
        for (int x = tid; x < cols; x += ENC_COMPUTE) ((T*)(sIn + y * PITCH))[x] = src[(size_t)y * a.nCols + x];
      namedBarSync(1, ENC_COMPUTE);
    }

    // ---- size: two threads per block (rows 0-3 / 4-7)
    for (int bb = tid >> 1; bb < TW; bb += ENC_COMPUTE / 2) {
      const int hf = tid & 1;
      const bool act = bb < nbk;
      const int w = act ? (BATCH ? 8 : min(8, a.nCols - (bx0 + bb) * 8)) : 0;
      const bool full = act && h == 8 && w == 8;
      bool hot = false;
      uint4 info = make_uint4(0, 0, 0, 0);
      uint32_t len = 0;
      if (hotType) {
        // hot path test for full 8x8 float blocks coded "bit-stuffed, <= 16 bits, offset as float/short/byte".  The thread with hf == 1
        // reads the two halves of a row in swapped order (bank-conflict free); min / max / the filters do not care about the order.
        const uint8_t* base = sIn + (hf * 4) * PITCH + bb * ROWB;
        float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000), t0 = 0.f;   // t0: NaN iff some value is NaN or +-Inf
        bool eq = false, ni = ((flagsSeen | myFlags) & FASTF_NOT_INT) != 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint4 A = *(const uint4*)(base + i * PITCH + hf * 16), B = *(const uint4*)(base + i * PITCH + 16 - hf * 16);
          const float a0 = __uint_as_float(A.x), a1 = __uint_as_float(A.y), a2 = __uint_as_float(A.z), a3 = __uint_as_float(A.w);
          const float b0 = __uint_as_float(B.x), b1 = __uint_as_float(B.y), b2 = __uint_as_float(B.z), b3 = __uint_as_float(B.w);
          mn = fminf(fminf(fminf(mn, a0), fminf(a1, a2)), fminf(fminf(a3, b0), fminf(fminf(b1, b2), b3)));
          mx = fmaxf(fmaxf(fmaxf(mx, a0), fmaxf(a1, a2)), fmaxf(fmaxf(a3, b0), fmaxf(fmaxf(b1, b2), b3)));
          t0 = __fmaf_rn(a0, 0.f, t0); t0 = __fmaf_rn(a1, 0.f, t0); t0 = __fmaf_rn(a2, 0.f, t0); t0 = __fmaf_rn(a3, 0.f, t0);
          t0 = __fmaf_rn(b0, 0.f, t0); t0 = __fmaf_rn(b1, 0.f, t0); t0 = __fmaf_rn(b2, 0.f, t0); t0 = __fmaf_rn(b3, 0.f, t0);
          // equal neighbours inside a 16-byte group: 48 of the block's 64 (value, predecessor) pairs.  No such pair => at most 16 equal
          // pairs => never a LUT candidate (needs more than 32, Lerc2.cpp:1794); otherwise the exact count is taken below.
          eq |= (a0 == a1) | (a1 == a2) | (a2 == a3) | (b0 == b1) | (b1 == b2) | (b2 == b3);
          if (!ni) {                                                  // all-integer test (Lerc.h:248): exact per thread, cheap once a fraction was seen
            ni = a0 != truncf(a0);
            if (!ni) ni = (a1 != truncf(a1)) | (a2 != truncf(a2)) | (a3 != truncf(a3)) | (b0 != truncf(b0)) | (b1 != truncf(b1)) | (b2 != truncf(b2)) | (b3 != truncf(b3));
          }
        }
        mn = fminf(mn, __shfl_xor_sync(FULL, mn, 1)); mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 1));
        unsigned pk = (t0 != t0 ? 1u : 0u) | (eq ? 2u : 0u);
        pk |= __shfl_xor_sync(FULL, pk, 1);
        const double zMn = (double)mn, zMx = (double)mx;
        const double mv = __dmul_rn(__dsub_rn(zMx, zMn), a.scale);
        const uint32_t me = roundToUInt(mv);
        const int nbh = bitLength(me);
        const bool lutMaybe = (pk & 2) && (zMx > __dadd_rn(zMn, a.maxZErr3));
        hot = full && !(pk & 1) && !(mv > (double)a.maxQ) && me > 0 && nbh <= 16 && !lutMaybe && !(mn == 0.f && mx == 0.f);
        if (hot) {
          // offset in the smallest type that holds it (Lerc2.h:457-542, float row)
          const bool isInt = mn == truncf(mn);
          const int tc = (isInt && mn >= 0.f && mn <= 255.f) ? 2 : ((isInt && mn >= -32768.f && mn <= 32767.f) ? 1 : 0);
          len = (uint32_t)(3 + (4 >> tc) + 8 * nbh);
          info.x = __float_as_uint(mn); info.y = (uint32_t)(nbh | (tc << 5) | (BEM_SIMPLE << 7) | TINFO_HOT);
          const uint32_t kmn = toKey(mn), kmx = toKey(mx);
          gMin = kmn < gMin ? (K)kmn : gMin; gMax = kmx > gMax ? (K)kmx : gMax;
          if (ni) myFlags |= FASTF_NOT_INT;
        }
      }
      if (!hot && act && hf == 0) {
        unsigned int fl = 0; K kmin, kmax;
        tileGenericChoice<T>(a, sIn + bb * ROWB, PITCH, h, w, info, len, fl, kmin, kmax);
        myFlags |= fl;
        gMin = kmin < gMin ? kmin : gMin; gMax = kmax > gMax ? kmax : gMax;
      }
      if (hf == 0) { sInfo[bb] = info; sOff[bb] = len; }
    }
    if (BATCH) {                                                     // the image's extremes and flags, from this tile
      K mn = gMin, mx = gMax;
      if constexpr (sizeof(K) == 4) { mn = __reduce_min_sync(FULL, mn); mx = __reduce_max_sync(FULL, mx); }
      else {
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) { const K omin = shflXorK<K>(mn, m), omax = shflXorK<K>(mx, m); mn = omin < mn ? omin : mn; mx = omax > mx ? omax : mx; }
      }
      const unsigned int fl = __reduce_or_sync(FULL, myFlags);
      if (lane == 0) {
        TileEncResult* r = &fb.imgRes[img];
        if (mx >= mn) { atomicMax(&r->negMinKey, ~(unsigned long long)mn); atomicMax(&r->maxKey, (unsigned long long)mx); }
        if (fl) atomicOr(&r->flags, fl);
      }
      gMin = keyMaxValue<K>(); gMax = 0; myFlags = 0;
    }
    namedBarSync(1, ENC_COMPUTE);

    // ---- scan (warp 0): block lengths -> exclusive byte offsets inside the tile, sOff[TW] = the tile's bytes; publish; post
    if (warp == 0) {
      constexpr int PER = TW / 32;
      uint32_t v[PER], sum = 0;
#pragma unroll
      for (int j = 0; j < PER; j++) { v[j] = sOff[lane * PER + j]; sum += v[j]; }
      uint32_t inc = sum;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) { const uint32_t o = __shfl_up_sync(FULL, inc, s); if (lane >= s) inc += o; }
      uint32_t run = inc - sum;
#pragma unroll
      for (int j = 0; j < PER; j++) { sOff[lane * PER + j] = run; run += v[j]; }
      if (lane == 31) {
        sOff[TW] = inc;
        if (BATCH) { fb.tileLen[tile] = inc; __threadfence(); }                 // (read by the control warps of the image's later tiles)
        lookbackPublish(st, a.groupAcc, tile, inc);                                     // for the other CTAs' look-backs
        sMailTile[k & 1] = tile; sMailBytes[k & 1] = inc;
        mbarArrive(&sBarScan[k & 1]);                                          // for this CTA's control warp
      }
    }
    namedBarSync(1, ENC_COMPUTE);
    const uint32_t tileBytes = sOff[TW];
    if (tid == 0) nextT = a.tileBegin + (int)atomicAdd(a.ticket, 1u);      // ticket of tile k + 1: in flight while this tile is packed
    // row 0 of the image against the coarser decimal grids (block row 0 only; behind the publication, so that nobody's look-back waits for it)
    if (!BATCH && isFlt && a.nRaise > 0 && tyT == 0) encRaiseRow0<T>(a, (const T*)sIn, min(TW * 8, a.nCols - bx0 * 8), tid, ENC_COMPUTE);

    if (tileBytes <= (uint32_t)C::STAGE_CAP) {
      // ---- the common case: one pass into this tile's staging image; it is flushed when the next tile has been packed
      packBlocks(stage, bx0, h, 0, nbk, 0u);
      if (tid == 0) sTileNext = nextT;
      namedBarSync(1, ENC_COMPUTE);                                  // image complete, sIn free
      const int tn = sTileNext;
      if (tid == 0 && vecOk && tn < a.tileEnd) issueLoad(tn);
      if (pendK >= 0) {
        mbarWait(&sBarOff[pendK & 1], (uint32_t)(pendK >> 1) & 1u);
        uint32_t* pst = (uint32_t*)(tileSmem + C::IN_BYTES + (pendK & 1) * C::STAGE_BYTES);
        flushPass(pst + 4, sOffS[pendK & 1], sRegS[pendK & 1], pendBytes, pendImg);
        namedBarSync(1, ENC_COMPUTE);
        for (int i = tid; i < (int)((pendBytes + 16 + 48 + 15) >> 4) && i < C::STAGE_BYTES / 16; i += ENC_COMPUTE) ((uint4*)pst)[i] = make_uint4(0, 0, 0, 0);
      }
      pendK = k; pendBytes = tileBytes; pendImg = img;
      tile = tn;
    } else {
      // ---- more than STAGE_CAP bytes (values wider than 16 bits, raw blocks): passes of at most STAGE_CAP bytes, each flushed at
      // once; the tile's offset is waited for here
      if (pendK >= 0) {                                              // first the waiting tile: its image is the other one
        mbarWait(&sBarOff[pendK & 1], (uint32_t)(pendK >> 1) & 1u);
        uint32_t* pst = (uint32_t*)(tileSmem + C::IN_BYTES + (pendK & 1) * C::STAGE_BYTES);
        flushPass(pst + 4, sOffS[pendK & 1], sRegS[pendK & 1], pendBytes, pendImg);
        namedBarSync(1, ENC_COMPUTE);
        for (int i = tid; i < C::STAGE_BYTES / 16; i += ENC_COMPUTE) ((uint4*)pst)[i] = make_uint4(0, 0, 0, 0);
        pendK = -1;
      }
      mbarWait(&sBarOff[k & 1], (uint32_t)(k >> 1) & 1u);
      const unsigned long long tileOff = sOffS[k & 1];
      const long long tileReg = sRegS[k & 1];
      int bLo = 0;
      uint32_t passBase = 0;
      while (bLo < nbk) {
        int bHi = nbk;
        if (tileBytes - passBase > (uint32_t)C::STAGE_CAP) {           // largest bHi with sOff[bHi] - passBase <= STAGE_CAP (a block is at most 513 bytes)
          int lo = bLo + 1, hi = nbk;
          while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sOff[mid] - passBase <= (uint32_t)C::STAGE_CAP) lo = mid; else hi = mid - 1; }
          bHi = lo;
        }
        const uint32_t passBytes = (bHi == nbk ? tileBytes : sOff[bHi]) - passBase;
        packBlocks(stage, bx0, h, bLo, bHi, passBase);
        namedBarSync(1, ENC_COMPUTE);
        flushPass(stage, tileOff + passBase, tileReg + (long long)passBase, passBytes, img);
        namedBarSync(1, ENC_COMPUTE);
        for (int i = tid; i < C::STAGE_BYTES / 16; i += ENC_COMPUTE) ((uint4*)(stage - 4))[i] = make_uint4(0, 0, 0, 0);
        bLo = bHi; passBase += passBytes;
        if (bLo < nbk) namedBarSync(1, ENC_COMPUTE);
      }
      if (tid == 0) sTileNext = nextT;
      namedBarSync(1, ENC_COMPUTE);                                  // image cleared, sIn free
      const int tn = sTileNext;
      if (tid == 0 && vecOk && tn < a.tileEnd) issueLoad(tn);
      tile = tn;
    }
  }
  // ---- the last tile's image; stop the control warp
  if (pendK >= 0) {
    mbarWait(&sBarOff[pendK & 1], (uint32_t)(pendK >> 1) & 1u);
    flushPass((uint32_t*)(tileSmem + C::IN_BYTES + (pendK & 1) * C::STAGE_BYTES) + 4, sOffS[pendK & 1], sRegS[pendK & 1], pendBytes, pendImg);
  }
  if (tid == 0) { sMailTile[k & 1] = -1; mbarArrive(&sBarScan[k & 1]); }
  if (BATCH) return;                                               // (per-image facts went out tile by tile)

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
  namedBarSync(1, ENC_COMPUTE);
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
  // ---- zero fill behind the blob (the API zero-fills the whole output buffer, Lerc.cpp:374): every CTA a slice, as soon as the stream's
  // length is known (all CTAs are resident: the grid is sized by occupancy)
  if (a.fillEnd) {
    if (tid == 0) { while (*(volatile unsigned int*)&a.res->totalReady == 0) __nanosleep(100); }
    namedBarSync(1, ENC_COMPUTE);
    const unsigned long long total = (unsigned long long)a.dataStart + *(volatile unsigned long long*)&a.res->totalBytes;
    uint8_t* from = a.blob + total;
    if (total <= a.blobCap && from < a.fillEnd) {
      const unsigned long long n = (unsigned long long)(a.fillEnd - from);
      const unsigned long long per = ((n + gridDim.x - 1) / gridDim.x + 15) & ~15ull;
      const unsigned long long lo = (unsigned long long)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
      if (lo < n) {
        uint8_t* p0 = from + lo; uint8_t* p1 = from + hi;
        uint8_t* q0 = (uint8_t*)(((uintptr_t)p0 + 15) & ~(uintptr_t)15); if (q0 > p1) q0 = p1;
        uint8_t* q1 = (uint8_t*)((uintptr_t)p1 & ~(uintptr_t)15); if (q1 < q0) q1 = q0;
        for (uint8_t* q = p0 + tid; q < q0; q += ENC_COMPUTE) *q = 0;
        for (uint4* q = (uint4*)q0 + tid; q < (uint4*)q1; q += ENC_COMPUTE) *q = make_uint4(0, 0, 0, 0);
        for (uint8_t* q = q1 + tid; q < p1; q += ENC_COMPUTE) *q = 0;
      }
    }
  }
}

}  // namespace lerc
