// lerc_tiles_decode.cuh -- tile batch decoder (included at the end of lerc_decode.cu).
//
// lerc_b200_decodeTiles (include/lerc_b200.h): nTiles standard Lerc2 blobs, one per tileRows x tileCols window of a
// one-band raster, are decoded into the raster in two launches instead of one lerc_decode call per tile:
//   k_tiles_parse    one thread per blob: header (Lerc2.cpp:762-917), mask byte count, ranges, flag byte.  Blobs that are
//                    "all valid, nDepth 1, 8x8 micro-block stream" (what the batch encoder and the reference write for
//                    ordinary tiles) are marked for the batch kernel; every other kind is left to the general band
//                    decoder, which decides exactly like the reference what is malformed.
//   k_tiles_blocks   one CTA per blob: Fletcher-32 of the blob (Lerc2.cpp:1012-1064); then the micro-block stream goes
//                    through shared memory window by window: one thread hops from block header to block header (a tile's
//                    stream is short, so the walk the reference does serially over the whole raster is serial only per
//                    tile here and the CTAs of other tiles hide it), all threads decode the blocks found, 8 lanes per
//                    block with the row decoder of lerc_decode_fast.cuh (ReadTile, Lerc2.cpp:2025-2230).
// Any anomaly (checksum, malformed unit, wrong integrity bits, stream too short) marks the blob for the general decoder.
#pragma once

namespace lerc {

enum { TILED_DONE = 0, TILED_FAST = 1, TILED_GENERAL = 2, TILED_CONST = 3 };   // CONST: header-only blob of a constant image, value in rec.zMax

struct TileDecRec { uint32_t status, streamOff, streamLen, version, checksum, blobSize; double invScale, zMax; };

struct TilesDecArgs {
  const uint8_t* blobs; const unsigned long long* offsets;     // [nImg + 1]
  int nImg, nImgX, imgCols, imgRows, rasterCols, rasterRows;
  TileDecRec* recs; void* data;
};

__device__ __forceinline__ uint32_t tdRd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
__device__ __forceinline__ double tdRdF64(const uint8_t* p) { return __longlong_as_double((long long)((unsigned long long)tdRd32(p) | ((unsigned long long)tdRd32(p + 4) << 32))); }

template <class T>
__global__ void k_tiles_parse(TilesDecArgs a) {
  constexpr int DT = PixelTraits<T>::code;
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= a.nImg) return;
  TileDecRec rec;
  rec.status = TILED_GENERAL; rec.streamOff = 0; rec.streamLen = 0; rec.version = 0; rec.checksum = 0; rec.blobSize = 0; rec.invScale = 0; rec.zMax = 0;
  const int iy = img / a.nImgX, ix = img - iy * a.nImgX;
  const int rows = min(a.imgRows, a.rasterRows - iy * a.imgRows), cols = min(a.imgCols, a.rasterCols - ix * a.imgCols);
  const unsigned long long o0 = a.offsets[img], o1 = a.offsets[img + 1];
  const unsigned long long avail = o1 > o0 ? o1 - o0 : 0;
  const uint8_t* p = a.blobs + o0;
  do {
    if (avail < 14 || p[0] != 'L' || p[1] != 'e' || p[2] != 'r' || p[3] != 'c' || p[4] != '2' || p[5] != ' ') break;
    const int version = (int)tdRd32(p + 6);
    if (version < 3 || version > 6) break;
    const int hb = version >= 6 ? 90 : (version >= 4 ? 66 : 62);
    if (avail < (unsigned long long)hb + 4) break;
    rec.version = (uint32_t)version; rec.checksum = tdRd32(p + 10);
    int q = 14;
    const int nRows = (int)tdRd32(p + q), nCols = (int)tdRd32(p + q + 4); q += 8;
    int nDepth = 1;
    if (version >= 4) { nDepth = (int)tdRd32(p + q); q += 4; }
    const int numValid = (int)tdRd32(p + q), mb = (int)tdRd32(p + q + 4), blobSize = (int)tdRd32(p + q + 8), dt = (int)tdRd32(p + q + 12); q += 16;
    int passNoData = 0;
    if (version >= 6) { passNoData = p[q + 4]; q += 8; }
    const double maxZErr = tdRdF64(p + q), zMin = tdRdF64(p + q + 8), zMax = tdRdF64(p + q + 16);
    if (nRows != rows || nCols != cols || nDepth != 1 || numValid != rows * cols || mb != 8 || dt != DT || passNoData) break;
    if (blobSize < hb + 4 || (unsigned long long)blobSize > avail) break;
    if (zMin == zMax) {                                                                  // constant image: header + empty mask, nothing else (Lerc2.cpp:410-413)
      if (blobSize == hb + 4 && tdRd32(p + hb) == 0) { rec.status = TILED_CONST; rec.blobSize = (uint32_t)blobSize; rec.zMax = zMin; }
      break;
    }
    if (!(maxZErr >= 0)) break;
    if ((dt == DT_Byte || dt == DT_Char) && maxZErr == 0.5) break;                       // an image-mode byte follows (Huffman)
    if (version >= 6 && dt >= DT_Float && maxZErr == 0) break;                           // likewise (lossless float)
    int pos = hb;
    if (tdRd32(p + pos) != 0) break;                                                     // all valid: no mask bytes (Lerc2.cpp:961-1008)
    pos += 4;
    if (version >= 4) {                                                                  // per-depth ranges (Lerc2.cpp:2643-2677)
      if (pos + 2 * (int)sizeof(T) > blobSize) break;
      T lo, hi; uint8_t tb[16];
      for (int i = 0; i < 2 * (int)sizeof(T); i++) tb[i] = p[pos + i];
      memcpy(&lo, tb, sizeof(T)); memcpy(&hi, tb + sizeof(T), sizeof(T));
      if ((double)lo == (double)hi) break;                                               // constant image
      pos += 2 * (int)sizeof(T);
    }
    if (pos + 1 > blobSize || p[pos] != 0) break;                                        // one sweep
    pos += 1;
    if (pos >= blobSize) break;
    rec.status = TILED_FAST; rec.streamOff = (uint32_t)pos; rec.streamLen = (uint32_t)(blobSize - pos); rec.blobSize = (uint32_t)blobSize;
    rec.invScale = __dmul_rn(2.0, maxZErr); rec.zMax = zMax;
  } while (false);
  a.recs[img] = rec;
}

// bytes of a block offset stored with type code tc for pixel type T (Lerc2.h:528-542), 0 when the code is not valid for T
template <class T> __device__ __forceinline__ int offsetSizeFromCode(int tc) {
  const int dtUsed = offsetTypeFromCode(PixelTraits<T>::code, tc);
  return dtUsed == DT_Undefined ? 0 : dtSize(dtUsed);
}

constexpr int TD_WIN = 4096;          // stream bytes walked per window
constexpr int TD_MAXW = 1024;         // blocks per window at most (flat areas: 1-3 bytes per block)

template <class T, int NT>
__global__ void __launch_bounds__(NT) k_tiles_blocks(TilesDecArgs a) {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  constexpr int LOOK = ((MAXU + 15) / 16) * 16 + 32;
  __shared__ __align__(16) uint8_t sBuf[TD_WIN + LOOK + 16];
  __shared__ uint16_t sPos[TD_MAXW + 2];
  __shared__ unsigned long long sA[NT / 32], sD[NT / 32];
  __shared__ int sN, sBad, sBlk, sTx, sTy;
  __shared__ uint32_t sCur;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = tid >> 3, r = tid & 7;
  T* data = (T*)a.data;
  const bool vecOk = (((long long)a.rasterCols * (long long)sizeof(T)) % 16 == 0) && ((a.imgCols * (int)sizeof(T)) % 16 == 0) && (((uintptr_t)data & 15) == 0);

  for (int img = blockIdx.x; img < a.nImg; img += gridDim.x) {
    const TileDecRec rec = a.recs[img];
    if (rec.status != TILED_FAST && rec.status != TILED_CONST) continue;      // uniform for the CTA
    const int iy = img / a.nImgX, ix = img - iy * a.nImgX;
    const int rows = min(a.imgRows, a.rasterRows - iy * a.imgRows), cols = min(a.imgCols, a.rasterCols - ix * a.imgCols);
    if (rec.status == TILED_CONST) {                                          // Lerc2::FillConstImage (Lerc2.cpp:2681-2721): every pixel = (T)zMin
      if (tid == 0) {
        const uint8_t* cb = a.blobs + a.offsets[img] + 14;
        const int len = (int)rec.blobSize - 14;                               // <= 80 header bytes behind the checksum field
        unsigned long long A = 0, D = 0;
        fletcherHostPartial(cb, 0, len, A, D);
        sBad = fletcherFinish(A, D, len) == rec.checksum ? 0 : 1;
      }
      __syncthreads();
      const bool ok = sBad == 0;
      if (ok) {
        const T v = (T)rec.zMax;
        for (int k = tid; k < rows * cols; k += NT) {
          const int rr = k / cols, cc = k - rr * cols;
          data[((long long)iy * a.imgRows + rr) * (long long)a.rasterCols + (long long)ix * a.imgCols + cc] = v;
        }
      }
      __syncthreads();                                                        // sBad is reused by the next blob
      if (tid == 0) a.recs[img].status = ok ? TILED_DONE : TILED_GENERAL;
      continue;
    }
    const int nTx = (cols + 7) / 8, nTy = (rows + 7) / 8, nBlocks = nTx * nTy;
    const uint8_t* blob = a.blobs + a.offsets[img];
    const int version = (int)rec.version;
    // ---- Fletcher-32 over blob bytes [14, blobSize): this pass also pulls the blob into L2 for the windows below
    {
      const uint8_t* region = blob + 14;
      const long long len = (long long)rec.blobSize - 14;
      const int dd = (int)((uintptr_t)region & 15);
      const uint4* g0 = (const uint4*)(region - dd);
      const long long nChunks = (len + dd + 15) >> 4;
      unsigned long long fa = 0, fd = 0;
      for (long long c = tid; c < nChunks; c += NT) {
        const uint4 x = __ldg(g0 + c);
        uint32_t o[4] = {x.x, x.y, x.z, x.w};
        const long long r0 = c * 16 - dd;
        if (r0 < 0 || r0 + 16 > len) {
#pragma unroll
          for (int j = 0; j < 16; j++) if (r0 + j < 0 || r0 + j >= len) o[j >> 2] &= ~(0xffu << (8 * (j & 3)));
        }
        const unsigned par = (unsigned)(r0 & 1);
        const uint32_t w0 = (uint32_t)((unsigned long long)(r0 + 16) >> 1) % 65535u;
        uint32_t S = 0, S1 = 0, prev = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
          const uint32_t cw = k < 4 ? o[k] : 0u;
          const uint32_t y = __funnelshift_l(prev, cw, par * 8);
          const uint32_t pw = __byte_perm(y, 0, 0x2301);
          const uint32_t wlo = pw & 0xffffu, whi = pw >> 16;
          S += wlo + whi; S1 += (uint32_t)(2 * k) * (wlo + whi) + whi;
          prev = cw;
        }
        fa += S; fd += (unsigned long long)(w0 + 65535u - 8u) * S + S1;
      }
      fa %= 65535ull; fd %= 65535ull;
#pragma unroll
      for (int m = 16; m; m >>= 1) { fa += __shfl_xor_sync(FULL, fa, m); fd += __shfl_xor_sync(FULL, fd, m); }
      if (lane == 0) { sA[warp] = fa; sD[warp] = fd; }
      if (tid == 0) { sBad = 0; sBlk = 0; sTx = 0; sTy = 0; sCur = 0; }
      __syncthreads();
      if (tid == 0) {
        unsigned long long A = 0, D = 0;
        for (int i = 0; i < NT / 32; i++) { A += sA[i]; D += sD[i]; }
        if (fletcherFinish(A, D, len) != rec.checksum) sBad = 1;
      }
      __syncthreads();
    }
    const uint8_t* stream = blob + rec.streamOff;
    const uint32_t streamLen = rec.streamLen;
    const size_t org = (size_t)iy * a.imgRows * (size_t)a.rasterCols + (size_t)ix * a.imgCols;

    while (!sBad && sBlk < nBlocks) {
      // ---- stage [cur, cur + TD_WIN + LOOK) of the stream: sBuf[d + i] = stream[cur + i]
      const uint32_t cur = sCur;
      const uint8_t* gcur = stream + cur;
      const int d = (int)((uintptr_t)gcur & 15);
      {
        const uint8_t* g0 = gcur - d;
        const long long availB = (long long)streamLen - (long long)cur + d;      // bytes of the blob's stream from g0
        for (int i = tid; i < (TD_WIN + LOOK + 16) / 16; i += NT) {
          uint4 x = make_uint4(0, 0, 0, 0);
          if ((long long)i * 16 < availB) x = __ldg((const uint4*)g0 + i);          // may read up to 15 bytes past the stream: inside the blob buffer + its 16-byte slack
          ((uint4*)sBuf)[i] = x;
        }
      }
      __syncthreads();
      // ---- one thread hops from header to header (ReadTiles, Lerc2.cpp:1672-1713)
      if (tid == 0) {
        uint32_t pos = cur;
        int blk = sBlk, tx = sTx, ty = sTy, n = 0, bad = 0;
        while (blk < nBlocks && pos < cur + TD_WIN && n < TD_MAXW) {
          if (pos >= streamLen) { bad = 1; break; }
          const int h = min(8, rows - ty * 8), w = min(8, cols - tx * 8);
          const uint8_t* p = sBuf + d + (pos - cur);
          const uint32_t flag = p[0];
          int len;
          // the common unit first: bit-stuffed, one-byte count, no LUT (flag | offset | numBits byte | count | packed values)
          const int oszQ = offsetSizeFromCode<T>((int)(flag >> 6));
          const uint32_t nbByte = p[1 + oszQ];
          if ((flag & 3) == 1 && !(version >= 5 && (flag & 4)) && oszQ > 0 && (nbByte & 0xe0) == 0x80) {
            if (p[2 + oszQ] != (uint32_t)(h * w)) { bad = 1; break; }
            len = 3 + oszQ + (int)(((uint32_t)(h * w) * (nbByte & 31) + 7) >> 3);
          } else {
            FdUnit u;
            if (!fdParse<T>(p, version, h * w, true, u) || u.len > MAXU) { bad = 1; break; }
            len = u.len;
          }
          if (fdPattern(flag, version) != (tx & (version >= 5 ? 14 : 15)) || (unsigned long long)pos + (unsigned)len > streamLen) { bad = 1; break; }
          sPos[n++] = (uint16_t)(pos - cur);
          pos += (uint32_t)len;
          blk++;
          if (++tx == nTx) { tx = 0; ty++; }
        }
        sPos[n] = (uint16_t)(pos - cur);
        sN = n; sCur = pos; if (bad) sBad = 1;
      }
      __syncthreads();
      // ---- all threads decode the blocks found: 8 lanes per block, lane r = block row r
      const int n = sN, blk0 = sBlk;
      const uint32_t* words = (const uint32_t*)sBuf;
      const uint8_t* sb = sBuf + d;
      bool bad = false;
      for (int i = g; i < n; i += NT / 8) {
        const int b = blk0 + i, ty = b / nTx, tx = b - ty * nTx;
        const int h = min(8, rows - ty * 8), w = min(8, cols - tx * 8);
        const int p = sPos[i];
        T out[8]; unsigned why = 0;
        const int len = fdDecodeBlockRow<T>(words, sb, d, p, version, tx & (version >= 5 ? 14 : 15), h * w, h, w, r, rec.invScale, rec.zMax, out, why);
        if (why || p + len != (int)sPos[i + 1]) { bad = true; continue; }
        if (r < h) fdStoreRow<T>(data + org + (size_t)(ty * 8 + r) * a.rasterCols + tx * 8, out, w, vecOk);
      }
      if (bad) sBad = 1;
      __syncthreads();
      if (tid == 0) { sBlk = blk0 + n; sTx = sBlk % nTx; sTy = sBlk / nTx; }
      __syncthreads();
    }
    if (tid == 0) a.recs[img].status = sBad ? TILED_GENERAL : TILED_DONE;
    __syncthreads();
  }
}

namespace {

// Decodes blob `img` alone with the general band decoder (= what lerc_decode does for it) into its window of the raster.
ErrCode decodeOneTile(Context* ctx, const TilesGeom& g, const uint8_t* dBlobs, const unsigned long long* hOffsets, long long img, void* dData) {
  const int rows = g.rowsOf(img), cols = g.colsOf(img);
  const size_t ts = (size_t)typeSize(g.dt);
  const int iy = (int)(img / g.nImgX), ix = (int)(img % g.nImgX);
  if (hOffsets[img + 1] <= hOffsets[img]) return Failed;
  const size_t avail = (size_t)(hOffsets[img + 1] - hOffsets[img]);
  ByteSource src; src.base = dBlobs + hOffsets[img]; src.size = avail; src.onDevice = true;
  BlobInfo li;
  ErrCode e = getBlobInfo(src, li, nullptr, nullptr, 0);
  if (e != Ok) return e;
  if (li.dt != g.dt || li.nDepth != 1 || li.nCols != cols || li.nRows != rows) return Failed;
  HeaderInfo hd; uint8_t head[96];
  const size_t take = std::min(avail, sizeof head);
  if (take < 14 || !src.fetch(0, take, head) || !readHeader(head, take, hd) || (size_t)hd.blobSize > avail) return Failed;
  if (li.nUsesNoDataValue) return Failed;
  const size_t arenaMark = ctx->arena.used, pinnedMark = ctx->pinnedUsed;
  void* dTile = ctx->arena.alloc((size_t)rows * cols * ts);
  BandMaskState ms;
  ms.dBits = (uint8_t*)ctx->arena.alloc(((size_t)rows * cols + 7) / 8);
  if (!dTile || !ms.dBits) return Failed;
  DecodeBandArgs a;
  a.dt = g.dt; a.nDepth = 1; a.nCols = cols; a.nRows = rows;
  a.dBlob = src.base; a.avail = avail; a.hd = hd; a.hBlob = nullptr; a.src = &src; a.srcOff = 0;
  a.dData = dTile; a.dValidBytes = nullptr;
  e = decodeBand(ctx, a, ms);
  if (e != Ok) return e;
  uint8_t* dst = (uint8_t*)dData + ((size_t)iy * g.tileRows * (size_t)g.nCols + (size_t)ix * g.tileCols) * ts;
  if (!cudaOk(cudaMemcpy2DAsync(dst, (size_t)g.nCols * ts, dTile, (size_t)cols * ts, (size_t)cols * ts, (size_t)rows, cudaMemcpyDeviceToDevice, ctx->stream), "tile scatter")) return Failed;
  if (!cudaOk(cudaStreamSynchronize(ctx->stream), "tile sync")) return Failed;
  if (ctx->arena.retired.empty()) ctx->arena.used = arenaMark;
  ctx->pinnedUsed = pinnedMark;
  return Ok;
}

template <class T>
ErrCode decodeTilesT(Context* ctx, const TilesGeom& g, const uint8_t* dBlobs, size_t blobBytes, const unsigned long long* hOffsets, void* dData) {
  const long long nImg = g.nImg();
  for (long long i = 0; i < nImg; i++) if (hOffsets[i + 1] < hOffsets[i] || hOffsets[i + 1] > blobBytes) return WrongParam;
  cudaStream_t st = ctx->stream;
  std::vector<TileDecRec> recs((size_t)nImg);
  const bool batch = !std::getenv("LERC_B200_NO_FAST") && nImg <= 0x7fffffffLL;
  if (batch) {
    unsigned long long* dOff = (unsigned long long*)ctx->arena.alloc((size_t)(nImg + 1) * 8);
    TileDecRec* dRecs = (TileDecRec*)ctx->arena.alloc((size_t)nImg * sizeof(TileDecRec));
    if (!dOff || !dRecs) return Failed;
    if (!cudaOk(cudaMemcpyAsync(dOff, hOffsets, (size_t)(nImg + 1) * 8, cudaMemcpyHostToDevice, st), "H2D offsets")) return Failed;
    TilesDecArgs a;
    a.blobs = dBlobs; a.offsets = dOff; a.nImg = (int)nImg; a.nImgX = g.nImgX; a.imgCols = g.tileCols; a.imgRows = g.tileRows;
    a.rasterCols = g.nCols; a.rasterRows = g.nRows; a.recs = dRecs; a.data = dData;
    LERC_LAUNCH(ctx, k_tiles_parse<T>, (unsigned)((nImg + 127) / 128), 128, 0, a);
    // CTA size: fewer threads per CTA = more CTAs (and header walkers) per SM; LERC_B200_TILES_NT = 32 (default: one warp per blob, no CTA-wide waiting on the header walk; measured 0.94 ms on 4096 tiles) | 64 (1.28 ms) | 128 (1.33 ms) | 256 (1.77 ms)
    static const int nt = [] { const char* e = std::getenv("LERC_B200_TILES_NT"); const int v = e ? std::atoi(e) : 32; return (v == 64 || v == 128 || v == 256) ? v : 32; }();
    auto launch = [&](auto kernel, int NT) {
      static DeviceInt occ;                         // all CTAs resident, each strides over the blobs: no tail wave
      int ctasPerSm = occ.get(ctx->device);
      if (!ctasPerSm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, kernel, NT, 0) != cudaSuccess || ctasPerSm < 1) ctasPerSm = 1;
        occ.set(ctx->device, ctasPerSm);
      }
      const long long grid = std::min<long long>(nImg, (long long)smCountOf(ctx->device) * ctasPerSm);
      LaunchScope scope_(ctx, "k_tiles_blocks<T>");
      kernel<<<(unsigned)grid, NT, 0, ctx->stream>>>(a); ctx->kernelLaunches++;
    };
    if (nt == 64) launch(k_tiles_blocks<T, 64>, 64); else if (nt == 128) launch(k_tiles_blocks<T, 128>, 128); else if (nt == 256) launch(k_tiles_blocks<T, 256>, 256); else launch(k_tiles_blocks<T, 32>, 32);
    if (!cudaOk(cudaMemcpyAsync(recs.data(), dRecs, (size_t)nImg * sizeof(TileDecRec), cudaMemcpyDeviceToHost, st), "D2H tile records")) return Failed;
    if (!cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
    if (!cudaOk(cudaGetLastError(), "decodeTiles")) return Failed;
  } else for (auto& r : recs) r.status = TILED_GENERAL;
  for (long long i = 0; i < nImg; i++) {
    if (recs[(size_t)i].status == TILED_DONE) { globalStats().fastPathDecodes++; continue; }
    const ErrCode e = decodeOneTile(ctx, g, dBlobs, hOffsets, i, dData);
    if (e != Ok) return e;
  }
  return Ok;
}

}  // namespace

ErrCode decodeTiles(Context* ctx, int dt, int nCols, int nRows, int tileCols, int tileRows, const uint8_t* dBlobs, size_t blobBytes,
                    const unsigned long long* hOffsets, void* dData) {
  TilesGeom g;
  g.dt = dt; g.nCols = nCols; g.nRows = nRows; g.tileCols = tileCols; g.tileRows = tileRows;
  g.nImgX = (nCols + tileCols - 1) / tileCols; g.nImgY = (nRows + tileRows - 1) / tileRows;
  switch (dt) {
    case DT_Char:   return decodeTilesT<int8_t>(ctx, g, dBlobs, blobBytes, hOffsets, dData);
    case DT_Byte:   return decodeTilesT<uint8_t>(ctx, g, dBlobs, blobBytes, hOffsets, dData);
    case DT_Short:  return decodeTilesT<int16_t>(ctx, g, dBlobs, blobBytes, hOffsets, dData);
    case DT_UShort: return decodeTilesT<uint16_t>(ctx, g, dBlobs, blobBytes, hOffsets, dData);
    case DT_Int:    return decodeTilesT<int32_t>(ctx, g, dBlobs, blobBytes, hOffsets, dData);
    case DT_UInt:   return decodeTilesT<uint32_t>(ctx, g, dBlobs, blobBytes, hOffsets, dData);
    case DT_Float:  return decodeTilesT<float>(ctx, g, dBlobs, blobBytes, hOffsets, dData);
    case DT_Double: return decodeTilesT<double>(ctx, g, dBlobs, blobBytes, hOffsets, dData);
    default: return WrongParam;
  }
}

}  // namespace lerc
