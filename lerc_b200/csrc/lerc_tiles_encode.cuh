// lerc_tiles_encode.cuh -- tile batch encoder (included at the end of lerc_encode.cu, inside namespace lerc).
//
// lerc_b200_encodeTiles (include/lerc_b200.h): a raster is cut into tileRows x tileCols images and EVERY image becomes
// its own standard Lerc2 blob -- byte for byte what lerc_encode returns for that pixel window alone -- written back to
// back into one output buffer, with an offset table.  This is the GPU form of the reference's callers that code
// rasters tile by tile (MRF / GeoTIFF tiles, BASELINE config 5: 65536^2 as 256^2 tiles; SURVEY.md 8(b) extension ii,
// 8(e)): one launch instead of one API call per tile.
//
//   k_try_raise_tiles   row 0 of every image against the coarser decimal grids (Lerc2.cpp:1233-1339), one warp per image
//   k_encode_fused<T, MINB, true>   the single-pass encoder of lerc_encode_fast.cuh over all images: tiles of 32 blocks
//                       never straddle an image, the look-back restarts at every image (offsets inside the blob) and
//                       a second look-back over whole blobs places each blob right behind its predecessor
//   k_tiles_finish      one thread per image: were the encoder's assumptions right for this image (same tests as
//                       encodeBandFast)?  If so header, ranges, flag byte and checksum are written on the device.
// Images whose assumptions failed (NaN, constant, all-integer floats, LUT candidates, 16x16 / one-sweep wins, ...)
// are encoded again by the general band encoder and the blobs behind them move up; so do all images when the
// pixel type or error bound is outside the fused encoder's domain.  The bytes are always those of lerc_encode.
#pragma once

namespace lerc {

template <class T>
__global__ void k_try_raise_tiles(const T* __restrict__ data, long long pitch, int nImg, int nImgX, int imgCols, int imgRows, int rasterCols,
                                  RaiseArgs ra, unsigned long long* __restrict__ maxBits /*[nImg][9]*/) {
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= nImg) return;
  const int iy = warp / nImgX, ix = warp - iy * nImgX;
  const int cols = min(imgCols, rasterCols - ix * imgCols);
  const T* row = data + (size_t)iy * imgRows * (size_t)pitch + (size_t)ix * imgCols;
  double m[9];
#pragma unroll
  for (int c = 0; c < 9; c++) m[c] = 0;
  for (int e = lane; e < cols; e += 32) {
    const double x = (double)row[e];
#pragma unroll
    for (int c = 0; c < 9; c++) {
      if (c < ra.n) {
        const double z = __dmul_rn(x, ra.fac[c]);
        const double dlt = fabs(__dsub_rn(floor(__dadd_rn(z, 0.5)), z));
        if (dlt > m[c]) m[c] = dlt;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 9; c++) {
    unsigned long long b = (unsigned long long)__double_as_longlong(m[c]);
    for (int s = 16; s; s >>= 1) { unsigned long long o = __shfl_xor_sync(FULL, b, s); b = o > b ? o : b; }
    if (lane == 0) maxBits[(size_t)warp * 9 + c] = b;
  }
}

// ALLINT: the only obstacle is that all-integer floats want max(0.5, floor(maxZErr)).  CONST: every pixel holds the same value; the blob is
// header + empty mask (TILE_CONST_BYTES), written at the start of the image's slot
// CONST_INT: a constant all-integer float image: its blob is the same for maxZErr and for max(0.5, floor(maxZErr))
enum { TILEST_OK = 0, TILEST_GENERAL = 1, TILEST_OVERFLOW = 2, TILEST_ALLINT = 3, TILEST_CONST = 4, TILEST_CONST_INT = 5 };
constexpr int TILE_CONST_BYTES = 90 + 4;

struct TileFinishArgs {
  const TileEncResult* res; const unsigned long long* imgState; const unsigned long long* raise;
  RaiseArgs ra; double raiseErr[9];                 // raiseErr[c]: the maxZError candidate c stands for (Lerc2.cpp:1242-1251)
  int nImg, nImgX, imgCols, imgRows, rasterCols, rasterRows, dataStart;
  double maxZErr;
  uint8_t* out; unsigned long long outCap;
  uint32_t* status;
  // after k_encode_tile<T, MINB, true>: the look-back words of all tiles (inclusive byte prefixes by then), tilesPerImg of them per image;
  // the end offset of every blob is written to imgStateOut.  nullptr after k_encode_fused (imgState holds the end offsets already).
  const unsigned long long* tileState; int tilesPerImg; unsigned long long* imgStateOut;
};

template <class T>
__global__ void k_tiles_finish(TileFinishArgs a) {
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  using K = typename PixelTraits<T>::Key;
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= a.nImg) return;
  TileEncResult r = a.res[img];
  const int iy = img / a.nImgX, ix = img - iy * a.nImgX;
  const int rows = min(a.imgRows, a.rasterRows - iy * a.imgRows), cols = min(a.imgCols, a.rasterCols - ix * a.imgCols);
  const long long nPix = (long long)rows * cols;
  constexpr unsigned long long VAL = (1ull << 62) - 1;
  unsigned long long start;
  if (a.tileState) {
    const long long first = (long long)img * a.tilesPerImg;
    const unsigned long long before = first == 0 ? 0ull : (a.tileState[first - 1] & VAL), behind = a.tileState[first + a.tilesPerImg - 1] & VAL;
    r.streamBytes = behind - before;
    start = before + (unsigned long long)img * (unsigned long long)a.dataStart;
    a.imgStateOut[img] = behind + (unsigned long long)(img + 1) * (unsigned long long)a.dataStart;
  } else start = img == 0 ? 0ull : (a.imgState[img - 1] & VAL);
  uint32_t st = TILEST_OK;
  // ---- the tests of encodeBandFast (lerc_encode.cu), per image
  if (r.flags & (FASTF_NAN | FASTF_LUT)) st = TILEST_GENERAL;
  const K minKey = (K)~r.negMinKey, maxKey = (K)r.maxKey;
  const T lo = fromKey<T>(minKey), hi = fromKey<T>(maxKey);
  const double zMin = (double)lo, zMax = (double)hi;
  if (zMin == zMax) st = TILEST_GENERAL;                                   // constant image: no stream at all
  uint8_t bIsInt = 0;
  bool allIntMismatch = false;
  const bool constImg = minKey == maxKey && !(r.flags & FASTF_NAN);       // one bit pattern everywhere (a mix of +0 and -0 is not)
  const bool hard = (r.flags & FASTF_NAN) || (zMin == zMax && !constImg);  // obstacles that do not depend on maxZError
  if (isFlt && (st == TILEST_OK || !hard)) {
    if (!constImg && ((zMin == 0 && __double_as_longlong(zMin) < 0) || (zMax == 0 && __double_as_longlong(zMax) >= 0))) st = TILEST_GENERAL;   // sign of a zero extreme
    bool allInt = !(r.flags & FASTF_NOT_INT);
    const double lim = sizeof(T) == 4 ? 8388608.0 : 9007199254740992.0;
    allInt = allInt && zMin >= -lim && zMin <= lim && zMax >= -lim && zMax <= lim;             // Lerc.cpp:1490-1500
    if (allInt) { const double f = floor(a.maxZErr); if ((f > 0.5 ? f : 0.5) != a.maxZErr) allIntMismatch = true; bIsInt = 1; }
    for (int c = 0; c < a.ra.n; c++) {                                     // PruneCandidates on row 0 (Lerc2.cpp:1322-1339)
      const double m = __longlong_as_double((long long)a.raise[(size_t)img * 9 + c]);
      if (!(__ddiv_rn(m, a.ra.fac[c]) > __dmul_rn(a.maxZErr, 0.5))) st = TILEST_GENERAL;     // a candidate survived: full scan needed
    }
  }
  const unsigned long long nData = r.streamBytes;
  const unsigned long long oneSweepBytes = sizeof(T) * (unsigned long long)nPix;
  if ((double)nData * 8 < (double)nPix * 1.5 && nData < 4 * oneSweepBytes && (rows > 8 || cols > 8)) st = TILEST_GENERAL;    // 16x16 retry (Lerc2.cpp:333-357)
  if (oneSweepBytes <= nData) st = TILEST_GENERAL;                         // one sweep raw wins (Lerc2.cpp:364-373)
  const unsigned long long total = (unsigned long long)a.dataStart + nData;
  // all-integer floats at a bound that is not max(0.5, floor(.)): the LUT / candidate / size tests above were made with the wrong bound;
  // unless something independent of the bound stands in the way, the image only needs the other bound
  if (allIntMismatch) st = (!hard && (constImg || !((zMin == 0 && __double_as_longlong(zMin) < 0) || (zMax == 0 && __double_as_longlong(zMax) >= 0)))) ? TILEST_ALLINT : TILEST_GENERAL;
  if ((r.flags & FASTF_OVERFLOW) || start + total > a.outCap) st = TILEST_OVERFLOW;
  // constant image (Lerc2.cpp:255-258): no ranges, no stream.  maxZError in the header: what TryRaiseMaxZError finds for the one
  // value there is (row 0 speaks for the whole image), unless the all-integer rule already fixed it
  double hdrMaxZErr = a.maxZErr;
  bool writeConst = false;
  if (constImg && st != TILEST_OVERFLOW) {
    if (isFlt && bIsInt) { const double f = floor(a.maxZErr); hdrMaxZErr = f > 0.5 ? f : 0.5; }          // Lerc.cpp:1490-1502
    if (isFlt && !bIsInt)
      for (int c = 0; c < a.ra.n; c++) {
        const double m = __longlong_as_double((long long)a.raise[(size_t)img * 9 + c]);
        if (!(__ddiv_rn(m, a.ra.fac[c]) > __dmul_rn(a.maxZErr, 0.5))) { hdrMaxZErr = a.raiseErr[c]; break; }
      }
    st = (isFlt && bIsInt) ? TILEST_CONST_INT : TILEST_CONST; writeConst = true;
  }
  a.status[img] = st;
  if (st != TILEST_OK && !writeConst) return;

  // ---- header (Lerc2.cpp:710-760, version 6), mask byte count 0, ranges, "not one sweep"; checksum over [14, total)
  uint8_t b[128];
  for (int i = 0; i < 128; i++) b[i] = 0;
  auto put32 = [&](int at, uint32_t v) { for (int i = 0; i < 4; i++) b[at + i] = (uint8_t)(v >> (8 * i)); };
  auto put64 = [&](int at, double d) { const unsigned long long v = (unsigned long long)__double_as_longlong(d); for (int i = 0; i < 8; i++) b[at + i] = (uint8_t)(v >> (8 * i)); };
  b[0] = 'L'; b[1] = 'e'; b[2] = 'r'; b[3] = 'c'; b[4] = '2'; b[5] = ' ';
  put32(6, 6u); put32(14, (uint32_t)rows); put32(18, (uint32_t)cols); put32(22, 1u); put32(26, (uint32_t)nPix); put32(30, 8u);
  put32(34, writeConst ? (uint32_t)TILE_CONST_BYTES : (uint32_t)total); put32(38, (uint32_t)PixelTraits<T>::code); put32(42, 0u);
  b[46] = 0; b[47] = bIsInt;
  put64(50, hdrMaxZErr); put64(58, zMin); put64(66, zMax);                  // noDataVal, noDataValOrig stay 0
  int p = 90 + 4;
  uint8_t* dst = a.out + start;
  if (writeConst) {
    unsigned long long A0 = 0, D0 = 0;
    fletcherHostPartial(b + 14, 0, (long long)p - 14, A0, D0);
    put32(10, fletcherFinish(A0, D0, (long long)p - 14));
    for (int i = 0; i < p; i++) dst[i] = b[i];
    return;
  }
  { uint8_t tmp[8]; memcpy(tmp, &lo, sizeof(T)); for (int i = 0; i < (int)sizeof(T); i++) b[p + i] = tmp[i]; p += (int)sizeof(T);
    memcpy(tmp, &hi, sizeof(T)); for (int i = 0; i < (int)sizeof(T); i++) b[p + i] = tmp[i]; p += (int)sizeof(T); }
  b[p++] = 0;
  unsigned long long A = r.fletA, D = r.fletD % 65535ull;
  fletcherHostPartial(b + 14, 0, (long long)p - 14, A, D);
  put32(10, fletcherFinish(A, D, (long long)total - 14));
  for (int i = 0; i < p; i++) dst[i] = b[i];
}

namespace {

// Encodes image `img` of the raster alone with the general band encoder (= what lerc_encode does for that window).
ErrCode encodeOneTile(Context* ctx, const TilesGeom& g, const void* dData, double maxZErr, long long img, uint8_t* dOut, size_t outCap, uint32_t& bytes) {
  const int rows = g.rowsOf(img), cols = g.colsOf(img);
  const size_t ts = (size_t)typeSize(g.dt);
  const int iy = (int)(img / g.nImgX), ix = (int)(img % g.nImgX);
  const size_t arenaMark = ctx->arena.used, pinnedMark = ctx->pinnedUsed;
  void* dTile = ctx->arena.alloc((size_t)rows * cols * ts);
  BandMaskState ms;
  ms.dPrevBits = (uint8_t*)ctx->arena.alloc(((size_t)rows * cols + 7) / 8);
  if (!dTile || !ms.dPrevBits) return Failed;
  const uint8_t* src = (const uint8_t*)dData + ((size_t)iy * g.tileRows * (size_t)g.nCols + (size_t)ix * g.tileCols) * ts;
  if (!cudaOk(cudaMemcpy2DAsync(dTile, (size_t)cols * ts, src, (size_t)g.nCols * ts, (size_t)cols * ts, (size_t)rows, cudaMemcpyDeviceToDevice, ctx->stream), "tile gather")) return Failed;
  EncodeBandArgs a;
  a.dt = g.dt; a.nDepth = 1; a.nCols = cols; a.nRows = rows; a.dData = dTile; a.dValidBytes = nullptr;
  a.maxZErr = maxZErr; a.iBand = 0; a.nBands = 1; a.nMasks = 0; a.anyMaskModified = false;
  a.dOut = dOut; a.outOffset = 0; a.outCapacity = outCap;
  const ErrCode e = encodeBand(ctx, a, ms, bytes);
  if (!cudaOk(cudaStreamSynchronize(ctx->stream), "tile sync")) return Failed;
  if (ctx->arena.retired.empty()) ctx->arena.used = arenaMark;
  ctx->pinnedUsed = pinnedMark;
  return e;
}

template <class T>
bool tilesFastEligible(const TilesGeom& g, double maxZErr) {
  if (sizeof(T) == 1 || std::getenv("LERC_B200_NO_FAST") || maxZErr == 777) return false;     // 777: bit-plane mode, decided per tile by the general encoder
  if (PixelTraits<T>::isFloat && !(maxZErr > 0)) return false;
  const long long nTxF = (g.tileCols + 7) / 8, nTyF = (g.tileRows + 7) / 8;
  const long long seg = ((nTxF + FAST_TB - 1) / FAST_TB) * nTyF;
  return seg * g.nImg() <= (1ll << 30) && (long long)g.tileCols * g.tileRows * (long long)sizeof(T) < (1ll << 30);
}

// The fused pass over all images.  hStatus[img] / hEnd[img] (end offset of blob img in dOut) on return.
template <class T>
ErrCode encodeTilesFast(Context* ctx, const TilesGeom& g, const void* dData, double maxZErrIn, uint8_t* dOut, size_t outCap,
                        std::vector<uint32_t>& hStatus, std::vector<unsigned long long>& hEnd) {
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  const double maxZErr = isFlt ? maxZErrIn : std::max(0.5, std::floor(maxZErrIn));            // Lerc2.cpp:219
  const long long nImg = g.nImg();
  cudaStream_t st = ctx->stream;
  const int nTxF = (g.tileCols + 7) / 8, nTyF = (g.tileRows + 7) / 8;
  // the TMA-staged persistent encoder (lerc_encode_tile.cuh, BATCH) when all images are equal, made of whole micro-blocks, and a tile of
  // TW blocks is a whole number of an image's block rows; else the older fused kernel
  constexpr int TWt = EncTile<T>::TW;
  const bool tma = g.nCols % g.tileCols == 0 && g.nRows % g.tileRows == 0 && g.tileCols % 8 == 0 && g.tileRows % 8 == 0 && nTxF <= TWt && TWt % nTxF == 0 &&
                   ((uintptr_t)dData & 15) == 0 && ((size_t)g.nCols * sizeof(T)) % 16 == 0 && !std::getenv("LERC_B200_TILES_OLD");
  if (std::getenv("LERC_B200_VERBOSE")) std::fprintf(stderr, "[lerc_b200] tile batch encoder: %s\n", tma ? "k_encode_tile (TMA, persistent)" : "k_encode_fused");
  const int rowsPerTile = tma ? TWt / nTxF : 1;
  const int segPerImg = tma ? (nTyF + rowsPerTile - 1) / rowsPerTile : ((nTxF + FAST_TB - 1) / FAST_TB) * nTyF;
  const long long nSeg = (long long)segPerImg * nImg;
  const int dataStart = headerBytes(6) + 4 + 2 * (int)sizeof(T) + 1;

  const size_t offRes = (size_t)nSeg * 8, offImg = offRes + (size_t)nImg * sizeof(TileEncResult), offRaise = offImg + (size_t)nImg * 8;
  const size_t nGroups = ((size_t)nSeg + 31) / 32;
  const size_t offStatus = offRaise + (size_t)nImg * 9 * 8, offGroups = (offStatus + (size_t)nImg * 4 + 15) / 16 * 16, offLen = offGroups + 2 * nGroups * 8;
  const size_t offMisc = offLen + (size_t)nSeg * 4, stateBytes = (offMisc + 15) / 16 * 16 + sizeof(FastEncResult) + 16;
  uint8_t* dState = (uint8_t*)ctx->arena.alloc(stateBytes);
  if (!dState) return Failed;
  cudaMemsetAsync(dState, 0, stateBytes, st);
  unsigned long long* dSegState = (unsigned long long*)dState;
  TileEncResult* dRes = (TileEncResult*)(dState + offRes);
  unsigned long long* dImgState = (unsigned long long*)(dState + offImg);
  unsigned long long* dRaise = (unsigned long long*)(dState + offRaise);
  uint32_t* dStatus = (uint32_t*)(dState + offStatus);

  RaiseArgs ra; ra.n = 0;
  double raiseErr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (isFlt) {
    static const double kErr[9] = {1, 0.5, 0.1, 0.05, 0.01, 0.005, 0.001, 0.0005, 0.0001};
    static const double kFac[9] = {1, 2, 10, 20, 100, 200, 1000, 2000, 10000};
    for (int i = 0; i < 9; i++) if (kErr[i] / 2 > maxZErr) { ra.fac[ra.n] = kFac[i]; raiseErr[ra.n] = kErr[i] / 2; ra.n++; }
    if (ra.n > 0) {
      ctx->forkSide();
      LERC_LAUNCH(ctx, k_try_raise_tiles<T>, (unsigned)((nImg + 7) / 8), 256, 0, (const T*)dData, (long long)g.nCols, (int)nImg, g.nImgX, g.tileCols, g.tileRows, g.nCols, ra, dRaise);
      ctx->backToMain();
    }
  }

  FastEncArgs fa; std::memset(&fa, 0, sizeof fa);
  fa.data = dData; fa.nRows = g.tileRows; fa.nCols = g.tileCols; fa.nTx = nTxF; fa.nTy = nTyF; fa.dt = PixelTraits<T>::code;
  fa.maxZErr = maxZErr; fa.scale = 1.0 / (2.0 * maxZErr); fa.maxZErr3 = 3.0 * maxZErr; fa.maxQ = fa.dt <= DT_UShort ? (1u << 15) - 1 : (1u << 30) - 1;   // Lerc2.h:685-703
  fa.intLossless = (!isFlt && maxZErr == 0.5) ? 1 : 0;
  fa.regionOff = (long long)dataStart - 14;
  fa.tileState = dSegState; fa.groupState = nullptr;
  FastBatchArgs fb;
  fb.imgCols = g.tileCols; fb.imgRows = g.tileRows; fb.nImgX = g.nImgX; fb.nImgY = g.nImgY; fb.rasterCols = g.nCols; fb.rasterRows = g.nRows;
  fb.segPerImg = segPerImg; fb.dataStart = dataStart; fb.pitch = g.nCols;
  fb.imgState = dImgState; fb.imgRes = dRes; fb.out = dOut; fb.outCap = outCap;
  fb.tilesPerImg = segPerImg; fb.tileLen = (uint32_t*)(dState + offLen);
  if (tma) {
    if (nSeg >= (1ll << 31) - 1) return Failed;
    FastEncResult* dMisc = (FastEncResult*)(dState + (offMisc + 15) / 16 * 16);
    fa.stream = dOut; fa.streamCap = outCap; fa.res = dMisc;
    fa.groupState = (unsigned long long*)(dState + offGroups); fa.groupAcc = fa.groupState + nGroups;
    fa.tileBegin = 0; fa.tileEnd = (int)nSeg; fa.ticket = &dMisc->ticket;
    constexpr size_t smem = (size_t)EncTile<T>::SMEM;
    static std::atomic<int> ctasPerSmOf[64];
    const int dv = ctx->device & 63;
    int ctasPerSm = ctasPerSmOf[dv].load(std::memory_order_relaxed);
    auto kernel = k_encode_tile<T, 3, true>;
    if (!ctasPerSm) {
      if (!cudaOk(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "encode tile smem")) return Failed;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, kernel, ENC_THREADS, smem) != cudaSuccess || ctasPerSm < 1) ctasPerSm = 1;
      ctasPerSmOf[dv].store(ctasPerSm, std::memory_order_relaxed);
    }
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const long long grid = std::min<long long>(nSeg, (long long)ctasPerSm * std::max(sms, 1));        // persistent CTAs, tiles by ticket
    { LaunchScope scope_(ctx, "k_encode_tile<T, tiles>"); kernel<<<(unsigned)grid, ENC_THREADS, smem, st>>>(fa, fb); ctx->kernelLaunches++; }
  } else {
    constexpr int MAXB = 1 + 64 * (int)sizeof(T);
    const size_t smem = (size_t)((FAST_TB * MAXB + 15) / 16 + 3) * 16 * 2 + 256 * 8 * sizeof(T);
    const int sms = smCountOf(ctx->device);
    static DeviceInt occ;
    int ctasPerSm = occ.get(ctx->device);
    auto kernel = k_encode_fused<T, 4, true>;
    if (!ctasPerSm) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, kernel, 256, smem) != cudaSuccess || ctasPerSm < 1) ctasPerSm = 1;
      occ.set(ctx->device, ctasPerSm);
    }
    const long long grid = std::min<long long>(nSeg, (long long)ctasPerSm * std::max(sms, 1));       // all CTAs co-resident (look-back)
    { LaunchScope scope_(ctx, "k_encode_fused<T, tiles>"); kernel<<<(unsigned)grid, 256, smem, ctx->stream>>>(fa, fb); ctx->kernelLaunches++; }
  }
  ctx->joinSide();
  TileFinishArgs ta;
  ta.res = dRes; ta.imgState = dImgState; ta.raise = dRaise; ta.ra = ra;
  for (int i = 0; i < 9; i++) ta.raiseErr[i] = raiseErr[i];
  ta.nImg = (int)nImg; ta.nImgX = g.nImgX; ta.imgCols = g.tileCols; ta.imgRows = g.tileRows; ta.rasterCols = g.nCols; ta.rasterRows = g.nRows;
  ta.dataStart = dataStart; ta.maxZErr = maxZErr; ta.out = dOut; ta.outCap = outCap; ta.status = dStatus;
  ta.tileState = tma ? dSegState : nullptr; ta.tilesPerImg = segPerImg; ta.imgStateOut = dImgState;
  LERC_LAUNCH(ctx, k_tiles_finish<T>, (unsigned)((nImg + 127) / 128), 128, 0, ta);

  hStatus.resize((size_t)nImg); hEnd.resize((size_t)nImg);
  if (!cudaOk(cudaMemcpyAsync(hStatus.data(), dStatus, (size_t)nImg * 4, cudaMemcpyDeviceToHost, st), "D2H tile status")) return Failed;
  if (!cudaOk(cudaMemcpyAsync(hEnd.data(), dImgState, (size_t)nImg * 8, cudaMemcpyDeviceToHost, st), "D2H tile offsets")) return Failed;
  if (!cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
  for (auto& v : hEnd) v &= (1ull << 62) - 1;
  return cudaOk(cudaGetLastError(), "encodeTilesFast") ? Ok : Failed;
}

template <class T>
ErrCode encodeTilesT(Context* ctx, const TilesGeom& g, const void* dData, double maxZErr, uint8_t* dOut, size_t outCap, unsigned long long* hOffsets) {
  const long long nImg = g.nImg();
  std::vector<uint32_t> status; std::vector<unsigned long long> end;
  bool fast = false;
  if constexpr (sizeof(T) > 1) fast = tilesFastEligible<T>(g, maxZErr);
  if constexpr (sizeof(T) > 1) if (fast) {
    ErrCode e = encodeTilesFast<T>(ctx, g, dData, maxZErr, dOut, outCap, status, end);
    if (e != Ok) return e;
    if (PixelTraits<T>::isFloat) {
      // integer-valued float rasters (elevation models stored as float): every tile wants max(0.5, floor(maxZErr)) (Lerc.cpp:1490-1502).
      // One more fused pass with that bound codes them all, instead of the general encoder tile by tile.
      // Tiles the first pass had to hand to the general encoder for another reason (constant, NaN, ...) stay handed over; the second
      // pass is only possible when the first one produced no blob worth keeping.
      long long nAllInt = 0, nKeep = 0;
      for (long long i = 0; i < nImg; i++) { nAllInt += status[(size_t)i] == TILEST_ALLINT; nKeep += status[(size_t)i] == TILEST_OK || status[(size_t)i] == TILEST_OVERFLOW || status[(size_t)i] == TILEST_CONST; }
      if (nAllInt > 0 && nKeep == 0) {
        const std::vector<uint32_t> first = status;
        e = encodeTilesFast<T>(ctx, g, dData, std::max(0.5, std::floor(maxZErr)), dOut, outCap, status, end);
        if (e != Ok) return e;
        for (long long i = 0; i < nImg; i++)
          if (first[(size_t)i] != TILEST_ALLINT && first[(size_t)i] != TILEST_CONST_INT && status[(size_t)i] != TILEST_OVERFLOW) status[(size_t)i] = TILEST_GENERAL;
      }
    }
    bool allOk = true, overflow = false;
    for (long long i = 0; i < nImg; i++) { allOk = allOk && status[(size_t)i] == TILEST_OK; overflow = overflow || status[(size_t)i] == TILEST_OVERFLOW; }
    if (allOk) {
      hOffsets[0] = 0;
      for (long long i = 0; i < nImg; i++) hOffsets[i + 1] = end[(size_t)i];
      globalStats().fastPathEncodes += (uint64_t)nImg;
      return Ok;
    }
    // An overflowing image leaves its successors unwritten: their bytes cannot be kept.  Everything goes through the
    // general encoder then (and very likely ends in BufferTooSmall there).
    if (overflow) fast = false;
  }
  // ---- repair: blobs of the images the fused pass got right are kept (moved up where needed), the others are
  // encoded by the general band encoder, in order, each right behind its predecessor
  uint8_t* dKeep = nullptr;
  if (fast) {
    const unsigned long long used = end[(size_t)nImg - 1];
    dKeep = (uint8_t*)ctx->arena.alloc((size_t)used + 16);
    if (!dKeep) return Failed;
    if (!cudaOk(cudaMemcpyAsync(dKeep, dOut, (size_t)used, cudaMemcpyDeviceToDevice, ctx->stream), "keep copy")) return Failed;
  }
  unsigned long long cursor = 0;
  hOffsets[0] = 0;
  long long i = 0;
  while (i < nImg) {
    if (fast && (status[(size_t)i] == TILEST_CONST || status[(size_t)i] == TILEST_CONST_INT)) {                       // constant image: its short blob sits at the start of its slot
      const unsigned long long from = i == 0 ? 0 : end[(size_t)i - 1];
      if (cursor + TILE_CONST_BYTES > outCap) return BufferTooSmall;
      if (!cudaOk(cudaMemcpyAsync(dOut + cursor, dKeep + from, TILE_CONST_BYTES, cudaMemcpyDeviceToDevice, ctx->stream), "blob move")) return Failed;
      cursor += TILE_CONST_BYTES;
      hOffsets[i + 1] = cursor;
      globalStats().fastPathEncodes += 1;
      i++;
    } else if (fast && status[(size_t)i] == TILEST_OK) {
      long long j = i;                                                    // run of good images: one copy
      while (j + 1 < nImg && status[(size_t)j + 1] == TILEST_OK) j++;
      const unsigned long long from = i == 0 ? 0 : end[(size_t)i - 1], to = end[(size_t)j];
      if (cursor + (to - from) > outCap) return BufferTooSmall;
      if (!cudaOk(cudaMemcpyAsync(dOut + cursor, dKeep + from, (size_t)(to - from), cudaMemcpyDeviceToDevice, ctx->stream), "blob move")) return Failed;
      for (long long k = i; k <= j; k++) hOffsets[k + 1] = cursor + (end[(size_t)k] - from);
      cursor += to - from;
      globalStats().fastPathEncodes += (uint64_t)(j - i + 1);
      i = j + 1;
    } else {
      uint32_t bytes = 0;
      const ErrCode e = encodeOneTile(ctx, g, dData, maxZErr, i, dOut + cursor, (size_t)(outCap - cursor), bytes);
      if (e != Ok) return e;
      cursor += bytes;
      hOffsets[i + 1] = cursor;
      i++;
    }
  }
  return cudaOk(cudaStreamSynchronize(ctx->stream), "sync") ? Ok : Failed;
}

}  // namespace

ErrCode encodeTiles(Context* ctx, int dt, int nCols, int nRows, int tileCols, int tileRows, const void* dData, double maxZErr,
                    uint8_t* dOut, size_t outCap, unsigned long long* hOffsets) {
  TilesGeom g;
  g.dt = dt; g.nCols = nCols; g.nRows = nRows; g.tileCols = tileCols; g.tileRows = tileRows;
  g.nImgX = (nCols + tileCols - 1) / tileCols; g.nImgY = (nRows + tileRows - 1) / tileRows;
  switch (dt) {
    case DT_Char:   return encodeTilesT<int8_t>(ctx, g, dData, maxZErr, dOut, outCap, hOffsets);
    case DT_Byte:   return encodeTilesT<uint8_t>(ctx, g, dData, maxZErr, dOut, outCap, hOffsets);
    case DT_Short:  return encodeTilesT<int16_t>(ctx, g, dData, maxZErr, dOut, outCap, hOffsets);
    case DT_UShort: return encodeTilesT<uint16_t>(ctx, g, dData, maxZErr, dOut, outCap, hOffsets);
    case DT_Int:    return encodeTilesT<int32_t>(ctx, g, dData, maxZErr, dOut, outCap, hOffsets);
    case DT_UInt:   return encodeTilesT<uint32_t>(ctx, g, dData, maxZErr, dOut, outCap, hOffsets);
    case DT_Float:  return encodeTilesT<float>(ctx, g, dData, maxZErr, dOut, outCap, hOffsets);
    case DT_Double: return encodeTilesT<double>(ctx, g, dData, maxZErr, dOut, outCap, hOffsets);
    default: return WrongParam;
  }
}

}  // namespace lerc
