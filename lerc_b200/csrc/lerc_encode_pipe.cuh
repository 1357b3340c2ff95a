// lerc_encode_pipe.cuh -- EXPERIMENTAL variant of the fused single-pass encoder (LERC_B200_ENC=pipe; default off, round-2 candidate).
//
// Why: ncu source view of k_encode_fused (profiles/README.md, "where the barrier stalls are"): 47 % of all warp samples are a barrier
// stall at the second __syncthreads of the tile loop -- seven warps wait there for warp 0, which does the decoupled look-back for
// the tile's byte offset after its own share of the packing.  This variant removes that dependency:
//   * the offset of tile i is looked up one tile LATER (while tile i + 1 is coded), when its predecessors' states have long been
//     published, so the look-back does not spin;
//   * EVERY warp does that look-back for itself (two dependent rounds through the 32-tile group aggregates of k_encode_fused's
//     LERC_B200_ENC_LB=group protocol: 64 cached 8-byte loads per warp and tile), so no warp waits for another one;
//   * three staging images rotate (tile i packs, tile i - 1 is flushed, tile i - 2's image is zeroed): ONE __syncthreads per tile
//     instead of two.
// Same bytes as k_encode_fused (checked on tools/cusim against the oracle, also under shuffled thread schedules); not yet timed on a
// B200.  BATCH = true is the tile batch form (lerc_tiles_encode.cuh): local and blob-level look-backs are both deferred and per warp,
// the per-image facts and checksum partials are handed over per warp with atomics, so the tile loop keeps its single barrier.
#pragma once

namespace lerc {

template <class T, int MINB, bool BATCH = false>
__global__ void __launch_bounds__(256, MINB) k_encode_pipe(FastEncArgs a, typename std::conditional<BATCH, FastBatchArgs, FastNoBatch>::type t) {
  using K = typename PixelTraits<T>::Key;
  constexpr bool isFlt = PixelTraits<T>::isFloat;
  constexpr int DT = PixelTraits<T>::code;
  constexpr int TB = FAST_TB;
  constexpr int MAXB = 1 + 64 * (int)sizeof(T);                 // longest block: raw (Lerc2.h:427)
  constexpr int NQ = (TB * MAXB + 15) / 16 + 3;                 // staging uint4s: 16 zero bytes | tile output | zero tail
  extern __shared__ __align__(16) uint32_t stageRaw[];          // THREE staging images (tile i packs, tile i - 1 is flushed, tile i - 2's is zeroed) | T sRow[256][8] (general path)
  T* sRow = (T*)(stageRaw + 3 * NQ * 4);
  __shared__ uint32_t sLen[2][TB];
  __shared__ unsigned long long sKMin[8], sKMax[8], sFA[8], sFD[8];
  __shared__ unsigned int sFlg[8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sb = lane >> 3, r = lane & 7;
  const unsigned int flagsSeen = BATCH ? 0u : *(volatile unsigned int*)&a.res->flags;
  const unsigned long long negMinSeen = BATCH ? 0ull : *(volatile unsigned long long*)&a.res->negMinKey, maxSeen = BATCH ? 0ull : *(volatile unsigned long long*)&a.res->maxKey;
  for (int i = tid; i < 3 * NQ; i += 256) ((uint4*)stageRaw)[i] = make_uint4(0, 0, 0, 0);

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
  uint32_t prevBytes = 0, prevPrevBytes = 0;
  int prevTile = -1, prevImg = 0, prevLocal = 0;

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

  // ---- deferred: global byte offset of a tile (two-round look-back through the group aggregates, done by EVERY warp on its own one
  // tile later, when the predecessors' states are long published -- no warp waits for another one) and its flush to HBM
  auto finishTile = [&](const int tile, const uint32_t tileBytes, uint32_t* const stage, const int img, const int local) {
    unsigned long long excl = 0, imgBase = 0;
    if (!BATCH)
    {
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
        if (isP) needGroups = false;                                    // an inclusive prefix inside the group: excl is already global
      }
      if (warp == 0 && l == 31 && needGroups && lane == 0) gs[g] = ST_A | (excl + tileBytes);    // this group's bytes (its 32 tiles are all sized)
      if (needGroups) {
        long long base = g - 1;
        for (;;) {
          const long long idx = base - lane;
          unsigned long long s = ST_P;                                  // virtual groups before 0: prefix 0
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
      if (warp == 0 && lane == 0) {
        if (tile > 0) st[tile] = ST_P | (excl + tileBytes);
        if (l == 31) gs[g] = ST_P | (excl + tileBytes);                 // inclusive prefix of the whole group
        if (tile == nTiles - 1) a.res->totalBytes = excl + tileBytes;
      }
    }
    else {
      // tile batch: the chain restarts at every image (offsets inside the blob); a second look-back over whole blobs places the blob
      const long long first = (long long)(tile - local);
      if (tile > first) {
        long long base = (long long)tile - 1;
        for (;;) {
          const long long idx = base - lane;
          unsigned long long s = ST_P;                                  // virtual tiles before the image's first: prefix 0
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
        if (warp == 0 && lane == 0) st[tile] = ST_P | (excl + tileBytes);
      }
      volatile unsigned long long* ist = t.imgState;
      const bool lastSeg = local == t.segPerImg - 1;
      const unsigned long long blobBytes = (unsigned long long)t.dataStart + excl + tileBytes;      // meaningful for the last segment only
      if (lastSeg && warp == 0 && lane == 0) { t.imgRes[img].streamBytes = excl + tileBytes; __threadfence(); ist[img] = (img == 0 ? ST_P : ST_A) | blobBytes; }
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
        if (lastSeg && warp == 0 && lane == 0) ist[img] = ST_P | (imgBase + blobBytes);
      }
    }
    const unsigned long long localOff = excl;                             // offset inside this blob's micro-block stream
    const unsigned long long tileOff = BATCH ? imgBase + (unsigned long long)t.dataStart + excl : excl;
    // ---- staging -> HBM in 16-byte chunks aligned to the GLOBAL address (the staging image is re-aligned with
    // funnel shifts), Fletcher-32 partial sums of the same words (bytes outside the tile are zero in the image);
    // every chunk read is zeroed again for the tile after next
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
    if (BATCH) {
      // the checksum partials of this tile belong to its image: every warp hands its share over (no CTA-wide reduction, no barrier)
      unsigned long long A = fa % 65535ull, D = fd % 65535ull;
#pragma unroll
      for (int m = 16; m; m >>= 1) { A += __shfl_xor_sync(FULL, A, m); D += __shfl_xor_sync(FULL, D, m); }
      unsigned int ov = __reduce_or_sync(FULL, overflow ? (unsigned int)FASTF_OVERFLOW : 0u);
      if (lane == 0) {
        TileEncResult* ir = t.imgRes + img;
        if (A | D) { atomicAdd(&ir->fletA, A); atomicAdd(&ir->fletD, D % 65535ull); }
        if (ov) atomicOr(&ir->flags, ov);
      }
      fa = 0; fd = 0; overflow = false;
    }
  };

  for (int it = 0; tile < nTiles; it++, tile += gridDim.x) {
    uint32_t* stage = stageRaw + (size_t)(it % 3) * NQ * 4 + 4;  // tile-local byte 0 of this tile's output image
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
    if (BATCH) {
      // image-global facts of this tile (min / max keys, flags) go to its image right away, per warp
      K wMin = gMin, wMax = gMax;
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) {
        const K omin = shflXorK<K>(wMin, m), omax = shflXorK<K>(wMax, m);
        wMin = omin < wMin ? omin : wMin; wMax = omax > wMax ? omax : wMax;
      }
      const unsigned int wf = __reduce_or_sync(FULL, myFlags);
      if (lane == 0) {
        TileEncResult* ir = t.imgRes + img;
        if (wMax >= wMin) { atomicMax(&ir->negMinKey, ~(unsigned long long)wMin); atomicMax(&ir->maxKey, (unsigned long long)wMax); }
        if (wf) atomicOr(&ir->flags, wf);
      }
      gMin = keyMaxValue<K>(); gMax = 0; myFlags = 0;
    }
    __syncthreads();                                             // the ONLY barrier per tile: block lengths visible; every warp has packed tile it - 1 and flushed tile it - 2
    if (it > 1) {  // zero what tile it - 2 used of its image (flushed by every warp before this barrier); tile it + 1 packs into it after the next barrier
      const int nz = (((int)prevPrevBytes + 15) >> 4) + 3;
      uint4* img = (uint4*)(stageRaw + (size_t)((it + 1) % 3) * NQ * 4);
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

    // ---- the previous tile: offset + flush (its image was completed by every warp before this iteration's barrier)
    if (it > 0) finishTile(prevTile, prevBytes, stageRaw + (size_t)((it - 1) % 3) * NQ * 4 + 4, prevImg, prevLocal);
    prevPrevBytes = prevBytes; prevBytes = tileBytes; prevTile = tile; prevImg = img; prevLocal = local;
    // ---- next tile
    cur = nxt; ty = nty; tx = ntx; h = nh; w = nw; act = nact;
    if (BATCH) { img = nimg; local = nlocal; org = norg; }
  }
  __syncthreads();                                               // the last tile is packed
  if (prevTile >= 0) finishTile(prevTile, prevBytes, stageRaw + (size_t)(((prevTile - (int)blockIdx.x) / (int)gridDim.x) % 3) * NQ * 4 + 4, prevImg, prevLocal);
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


}  // namespace lerc
