// lerc_decode_fast.cuh -- single-kernel parallel decoder of a Lerc2 micro-block stream for the common raster
// shape: every pixel valid, nDepth == 1, 8x8 micro-blocks (included by lerc_decode.cu).
//
// The stream has no index: a block's length is only known from its own header bytes (ReadTile, Lerc2.cpp:2025-2230;
// BitStuffer2::Decode, BitStuffer2.cpp:159-258), so the reference walks it serially.  Here one persistent
// cooperative kernel (one CTA per SM) finds the block boundaries speculatively and decodes:
//   phase 1  the stream is cut into 4 KB sub-chunks (one warp each, staged in shared memory).  Every byte position
//            in the first MAXU bytes of a sub-chunk that parses as a block header is a candidate entry; candidates
//            hop from header to header (32 of them per warp step) and die on the first malformed header or broken
//            integrity-bit sequence (Lerc2.cpp:2045).  Survivors reach the end of the sub-chunk: (entry, exit, #blocks).
//            Wrong candidates either die or merge into the true chain, so only a handful survive.
//   phase 1b the sub-chunk maps of a CTA's region are composed in shared memory -> region map; grid barrier.
//   phase 2  every CTA composes the (<= 148) region maps from stream position 0 up to its own region.
//   phase 3  true (entry position, first block index) of every sub-chunk of the region.
//   phase 4  8 lanes per micro-block, one lane per block row: unpack numBits-wide values with funnel shifts,
//            z = offset + q * 2 maxZError in fp64 without contraction, min(z, zMax), cast, 128-bit stores.
//            Every sub-chunk checks that its walk ends exactly where the next one starts, so the result is the
//            serial parse or an error flag -- never a silently different parse.
// Streams the speculation cannot handle (partial raw blocks, units longer than MAXU, too many surviving
// candidates) raise DECF_FALLBACK and the host runs the general decoder (decodeBandT).
#pragma once
#include <cooperative_groups.h>

namespace lerc {

enum { DECF_FALLBACK = 8 };
constexpr int FD_SUB = 4096;          // sub-chunk bytes
constexpr int FD_ENT = 8;             // surviving entry candidates kept per sub-chunk / region
constexpr int FD_WARPS = 16;
constexpr uint32_t FD_DEAD = 0xffffffffu;
constexpr int FD_WK = 576;            // walker slots per warp (>= longest unit + 31)

struct FdEntry { uint32_t entry, exit, count; };

struct FastDecArgs {
  const uint8_t* stream; unsigned long long streamLen;
  int nRows, nCols, nTx, nTy, dt, version;
  double invScale, zMax;                 // 2 * maxZError ; header zMax
  void* data;
  int nSub, subPerReg, nReg, maxU;       // maxU = 1 + 64 * sizeof(T)
  FdEntry* regTab; int* regN;            // [nReg][FD_ENT], [nReg]
  unsigned int* barrier; int* status;
};

// ---- unit header ---------------------------------------------------------------------------------
struct FdUnit { int mode, tc, osz, nb, lut, n, nLut, nbIdx, pay, lutPay, len; };

// Parses the unit whose first byte is p[0] (>= 16 readable bytes).  cells = pixels of the block (every pixel is
// valid here), so raw blocks hold cells values and bit-stuffed blocks must hold exactly cells values
// (Lerc2.cpp:2148).  Returns false when the reference's ReadTile / BitStuffer2::Decode would fail.
template <class T>
__device__ __forceinline__ bool fdParse(const uint8_t* __restrict__ p, int version, int cells, bool exact, FdUnit& u) {
  constexpr int DT = PixelTraits<T>::code;
  const unsigned flag = p[0];
  u.mode = flag & 3; u.tc = flag >> 6; u.osz = 0; u.nb = 0; u.lut = 0; u.n = 0; u.nLut = 0; u.nbIdx = 0; u.pay = 1; u.lutPay = 0;
  if (version >= 5 && (flag & 4)) return false;                       // depth-delta flag needs a previous depth (Lerc2.cpp:2045)
  if (u.mode == 2) { u.len = 1; return true; }
  if (u.mode == 0) { u.len = 1 + cells * (int)sizeof(T); return true; }
  const int dtUsed = offsetTypeFromCode(DT, u.tc);
  if (dtUsed == DT_Undefined) return false;
  u.osz = dtSize(dtUsed);
  if (u.mode == 3) { u.len = 1 + u.osz; return true; }
  const unsigned b = p[1 + u.osz], code = b >> 6;
  u.nb = b & 31; u.lut = (b >> 5) & 1;
  const int cb = code == 0 ? 4 : 3 - (int)code;
  if (cb <= 0) return false;
  unsigned n = p[2 + u.osz];
  if (cb >= 2) n |= (unsigned)p[3 + u.osz] << 8;
  if (cb == 4) n |= ((unsigned)p[4 + u.osz] << 16) | ((unsigned)p[5 + u.osz] << 24);
  if (exact ? (n != (unsigned)cells) : (n == 0 || n > (unsigned)cells)) return false;   // speculative walks do not know the block's size yet
  u.n = (int)n;
  int len = 2 + u.osz + cb;
  if (!u.lut) { u.pay = len; len += (int)packedBytes(n, u.nb); }
  else {
    if (u.nb == 0) return false;
    u.nLut = (int)p[len] - 1;
    if (u.nLut < 1) return false;
    len += 1; u.lutPay = len; len += (int)packedBytes((uint32_t)u.nLut, u.nb);
    u.nbIdx = bitLength((uint32_t)u.nLut);
    u.pay = len; len += (int)packedBytes(n, u.nbIdx);
  }
  u.len = len;
  return true;
}

// integrity bits of a block header (Lerc2.cpp:2045): (j0 >> 3) & 15, only bits 1..3 of it from version 5 on
__device__ __forceinline__ int fdPattern(unsigned flag, int version) { return (int)((flag >> 2) & (version >= 5 ? 14u : 15u)); }
// may a block with pattern b follow a block with pattern a in stream order?  (tx -> tx + 1, or a new block row)
__device__ __forceinline__ bool fdFollows(int a, int b, int version) {
  if (b == 0) return true;
  return version >= 5 ? (b == a || b == ((a + 2) & 14)) : (b == ((a + 1) & 15));
}

// bits [bit, bit + nb) of a little-endian bit stream in shared memory (byte pointer, any alignment)
__device__ __forceinline__ uint32_t fdExtract(const uint8_t* __restrict__ base, uint32_t bit, int nb) {
  const uint32_t byte = bit >> 3;
  const uint32_t* w = (const uint32_t*)((uintptr_t)(base + byte) & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(((uintptr_t)(base + byte) & 3) * 8 + (bit & 7));
  const unsigned long long x = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
  uint32_t v = (uint32_t)(x >> sh);
  if (sh + nb > 64) v |= w[2] << (64 - sh);
  return nb >= 32 ? v : (v & ((1u << nb) - 1));
}

// software grid barrier (all CTAs are co-resident: cooperative launch)
__device__ __forceinline__ void fdGridBarrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*(volatile unsigned int*)counter < target) { }
    __threadfence();
  }
  __syncthreads();
}

template <class T> __device__ __forceinline__ T fdCast(double z, double zMax) { const double v = z < zMax ? z : zMax; return (T)v; }   // Lerc2.cpp:2160

// ---- the kernel --------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(FD_WARPS * 32, 1) k_decode_fused(FastDecArgs a) {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  constexpr int BUFB = ((FD_SUB + MAXU + 64 + 15) / 16) * 16;       // per-warp staging of one sub-chunk (+ look-ahead)
  extern __shared__ __align__(16) uint8_t smem[];
  // layout: [FD_WARPS][BUFB] stream staging | FdEntry sTab[subPerReg][FD_ENT] | uint8 sTabN[subPerReg] | FdEntry sReg[nReg][FD_ENT] | int sRegN[nReg]
  //         | uint32 sTrue[subPerReg][2] | uint16 walker scratch [FD_WARPS][3][320]
  uint8_t* sp = smem;
  uint8_t* bufAll = sp; sp += (size_t)FD_WARPS * BUFB;
  FdEntry* sTab = (FdEntry*)sp; sp += (size_t)a.subPerReg * FD_ENT * sizeof(FdEntry);
  FdEntry* sReg = (FdEntry*)sp; sp += (size_t)a.nReg * FD_ENT * sizeof(FdEntry);
  uint32_t* sTrue = (uint32_t*)sp; sp += (size_t)(a.subPerReg + 1) * 2 * 4;
  int* sRegN = (int*)sp; sp += (size_t)a.nReg * 4;
  uint16_t* wkAll = (uint16_t*)sp; sp += (size_t)FD_WARPS * 3 * FD_WK * 2;
  uint8_t* sTabN = sp;
  __shared__ int sBad;
  __shared__ uint32_t sRegEntry[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int reg = blockIdx.x;
  const int sub0 = reg * a.subPerReg, sub1 = min(a.nSub, sub0 + a.subPerReg), nLocal = max(0, sub1 - sub0);
  const int nBlocks = a.nTx * a.nTy;
  const int version = a.version;
  uint8_t* buf = bufAll + (size_t)warp * BUFB;
  uint16_t* wPos = wkAll + (size_t)warp * 3 * FD_WK, *wEnt = wPos + FD_WK, *wCnt = wEnt + FD_WK;
  if (tid == 0) sBad = 0;
  __syncthreads();

  // stage sub-chunk s of the stream: buf[d + i] = stream[s * FD_SUB + i] for i in [0, FD_SUB + MAXU + 16), d = global misalignment
  auto stageSub = [&](int s) -> int {
    const unsigned long long start = (unsigned long long)s * FD_SUB;
    const uint8_t* g = a.stream + start;
    const int d = (int)((uintptr_t)g & 15);
    const uint8_t* g0 = g - d;
    const long long avail = (long long)(a.streamLen - start) + d;              // bytes from g0 that belong to the stream
    for (int i = lane; i < BUFB / 16; i += 32) {
      uint4 x = make_uint4(0, 0, 0, 0);
      if ((long long)i * 16 < avail) x = __ldg((const uint4*)g0 + i);
      ((uint4*)buf)[i] = x;
    }
    __syncwarp();
    return d;
  };

  // ================= phase 1: candidate walks per sub-chunk =========================================
  for (int ls = warp; ls < nLocal; ls += FD_WARPS) {
    const int s = sub0 + ls;
    const int d = stageSub(s);
    const uint8_t* sb = buf + d;                                               // sb[i] = stream[s * FD_SUB + i]
    const long long left = (long long)(a.streamLen - (unsigned long long)s * FD_SUB);   // stream bytes from the sub-chunk start
    const int subEnd = (int)min((long long)FD_SUB, left);                      // walkers stop once they reach subEnd
    // ---- stage A: every position of the head window that parses as a header becomes a walker
    int nW = 0;
    const int headEnd = s == 0 ? 1 : min(MAXU, subEnd);                        // sub-chunk 0 starts with block 0 at position 0
    for (int base = 0; base < headEnd; base += 32) {
      const int p = base + lane;
      bool ok = false; FdUnit u; int pat = 0;
      if (p < headEnd) {
        ok = fdParse<T>(sb + p, version, 64, false, u); pat = fdPattern(sb[p], version);
        if (ok && (long long)p + u.len > left) { if (u.mode == 0) u.len = (int)(left - p); else ok = false; }
      }
      const unsigned m = __ballot_sync(FULL, ok);
      if (ok) { const int i = nW + __popc(m & ((1u << lane) - 1)); wPos[i] = (uint16_t)(p + u.len); wEnt[i] = (uint16_t)p; wCnt[i] = (uint16_t)(1 | (pat << 12)); }
      nW += __popc(m);
    }
    __syncwarp();
    // ---- stage B: hop until every walker has left the sub-chunk or died; in-place stable compaction.
    // Finished walkers become results: lane i holds result i; when more than 32 finish the lowest entries stay.
    uint32_t rEnt = FD_DEAD, rExit = 0, rCnt = 0;
    int nRes = 0;
    while (nW > 0) {
      int nNew = 0;
      for (int base = 0; base < nW; base += 32) {
        const int i = base + lane;
        bool alive = false, done = false; int pos = 0, ent = 0, cnt = 0, pat = 0;
        if (i < nW) { pos = wPos[i]; ent = wEnt[i]; cnt = wCnt[i] & 0xfff; pat = wCnt[i] >> 12; alive = true; }
        if (alive && pos >= subEnd) { done = true; alive = false; }
        if (alive) {
          FdUnit u;
          const int np = fdPattern(sb[pos], version);
          const bool okHdr = fdParse<T>(sb + pos, version, 64, false, u) && fdFollows(pat, np, version) && cnt < 0xfff;
          if (okHdr && (long long)pos + u.len <= left) { pos += u.len; cnt++; pat = np; }
          else if (okHdr && u.mode == 0) { pos = (int)left; cnt++; pat = np; }   // a raw block of a partial tile at the very end is shorter than assumed; phase 4 checks
          else alive = false;
        }
        __syncwarp();
        const unsigned md = __ballot_sync(FULL, done);
        for (unsigned mm = md; mm; mm &= mm - 1) {
          const int src = __ffs(mm) - 1;
          const uint32_t e = __shfl_sync(FULL, (uint32_t)ent, src), x = __shfl_sync(FULL, (uint32_t)pos, src), c = __shfl_sync(FULL, (uint32_t)cnt, src);
          if (nRes < 32) { if (lane == nRes) { rEnt = e; rExit = x; rCnt = c; } nRes++; }
          else {
            const uint32_t mx = __reduce_max_sync(FULL, rEnt);
            const unsigned who = __ballot_sync(FULL, rEnt == mx);
            if (e < mx && lane == __ffs(who) - 1) { rEnt = e; rExit = x; rCnt = c; }
          }
        }
        const unsigned ma = __ballot_sync(FULL, alive);
        if (alive) { const int j = nNew + __popc(ma & ((1u << lane) - 1)); wPos[j] = (uint16_t)pos; wEnt[j] = (uint16_t)ent; wCnt[j] = (uint16_t)(cnt | (pat << 12)); }
        nNew += __popc(ma);
        __syncwarp();
      }
      nW = nNew;
    }
    // keep the FD_ENT lowest entries (positions relative to the stream start)
    int slot = lane;
    if (nRes > FD_ENT) {
      int rank = 0;
      for (int j = 0; j < 32; j++) { const uint32_t o = __shfl_sync(FULL, rEnt, j); rank += (o < rEnt) ? 1 : 0; }
      slot = rEnt == FD_DEAD ? 32 : rank;
    } else if (lane >= nRes) slot = 32;
    for (int e = lane; e < FD_ENT; e += 32) sTab[(size_t)ls * FD_ENT + e].entry = FD_DEAD;
    __syncwarp();
    if (slot < FD_ENT) { FdEntry e; e.entry = rEnt + (uint32_t)s * FD_SUB; e.exit = rExit + (uint32_t)s * FD_SUB; e.count = rCnt; sTab[(size_t)ls * FD_ENT + slot] = e; }
    if (lane == 0) sTabN[ls] = (uint8_t)min(nRes, FD_ENT);
    __syncwarp();
  }
  __syncthreads();

  // ================= phase 1b: compose the sub-chunk maps of this region ===============================
  if (warp == 0) {
    FdEntry cur; cur.entry = FD_DEAD; cur.exit = 0; cur.count = 0;
    const int n0 = nLocal > 0 ? sTabN[0] : 0;
    if (lane < n0) cur = sTab[lane];
    for (int ls = 1; ls < nLocal; ls++) {
      if (cur.entry != FD_DEAD) {
        bool found = false;
        const int n = sTabN[ls];
        for (int e = 0; e < n; e++) {
          const FdEntry t = sTab[(size_t)ls * FD_ENT + e];
          if (t.entry == cur.exit) { cur.exit = t.exit; cur.count += t.count; found = true; break; }
        }
        if (!found) cur.entry = FD_DEAD;
      }
    }
    if (lane < FD_ENT) a.regTab[(size_t)reg * FD_ENT + lane] = cur;
    if (lane == 0) a.regN[reg] = n0;
  }
  fdGridBarrier(a.barrier, gridDim.x);

  // ================= phase 2: true entry of this region =============================================
  for (int i = tid; i < a.nReg * FD_ENT; i += blockDim.x) sReg[i] = a.regTab[i];
  for (int i = tid; i < a.nReg; i += blockDim.x) sRegN[i] = a.regN[i];
  __syncthreads();
  if (warp == 0) {
    uint32_t pos = 0, blk = 0; bool bad = false;
    for (int rg = 0; rg < reg && !bad; rg++) {
      if (blk >= (uint32_t)nBlocks) break;
      const FdEntry t = lane < sRegN[rg] ? sReg[(size_t)rg * FD_ENT + lane] : FdEntry{FD_DEAD, 0, 0};
      const unsigned m = __ballot_sync(FULL, t.entry == pos && t.entry != FD_DEAD);
      if (!m) { bad = true; break; }
      const int src = __ffs(m) - 1;
      pos = __shfl_sync(FULL, t.exit, src); blk += __shfl_sync(FULL, t.count, src);
    }
    if (lane == 0) { sRegEntry[0] = bad ? FD_DEAD : pos; sRegEntry[1] = blk; if (bad) sBad = 1; }
  }
  __syncthreads();

  // ================= phase 3: true entry of every sub-chunk of the region ==============================
  if (warp == 0 && lane == 0) {
    uint32_t pos = sRegEntry[0], blk = sRegEntry[1];
    for (int ls = 0; ls <= nLocal; ls++) {
      sTrue[2 * ls] = pos; sTrue[2 * ls + 1] = blk;
      if (ls == nLocal || pos == FD_DEAD) { if (pos == FD_DEAD) for (int k = ls; k <= nLocal; k++) { sTrue[2 * k] = FD_DEAD; sTrue[2 * k + 1] = blk; } break; }
      if (blk >= (uint32_t)nBlocks) { for (int k = ls; k <= nLocal; k++) { sTrue[2 * k] = FD_DEAD - 1; sTrue[2 * k + 1] = blk; } break; }   // past the last block: nothing to decode
      bool found = false;
      const int n = sTabN[ls];
      for (int e = 0; e < n; e++) {
        const FdEntry t = sTab[(size_t)ls * FD_ENT + e];
        if (t.entry == pos) { pos = t.exit; blk += t.count; found = true; break; }
      }
      if (!found) { sBad = 1; pos = FD_DEAD; }
    }
  }
  __syncthreads();

  // ================= phase 4: decode ================================================================
  T* data = (T*)a.data;
  const bool vecOk = ((a.nCols * (int)sizeof(T)) % 16 == 0) && (((uintptr_t)data & 15) == 0);
  const int sbI = lane >> 3, r = lane & 7;
  bool bad = false, fallback = false;
  for (int ls = warp; ls < nLocal; ls += FD_WARPS) {
    const uint32_t pos0 = sTrue[2 * ls], blk0 = sTrue[2 * ls + 1];
    if (pos0 >= FD_DEAD - 1) continue;                                         // dead chain (reported through sBad) or past the end
    const int s = sub0 + ls;
    const uint32_t expectExit = sTrue[2 * ls + 2], expectBlk = sTrue[2 * ls + 3];
    const int d = stageSub(s);
    const uint8_t* sb = buf + d;
    const long long left = (long long)(a.streamLen - (unsigned long long)s * FD_SUB);
    const int subEnd = (int)min((long long)FD_SUB, left);
    int pos = (int)(pos0 - (uint32_t)s * FD_SUB);
    uint32_t blk = blk0;
    while (pos < subEnd && blk < (uint32_t)nBlocks && !bad && !fallback) {
      // ---- walk up to 4 units; group g decodes the g-th of them
      int myPos = -1, myBlk = 0;
#pragma unroll
      for (int g = 0; g < 4; g++) {
        if (pos < subEnd && blk < (uint32_t)nBlocks && !bad && !fallback) {
          const int ty = (int)blk / a.nTx, tx = (int)blk - ty * a.nTx;
          const int h = min(8, a.nRows - ty * 8), w = min(8, a.nCols - tx * 8);
          FdUnit u;
          if (!fdParse<T>(sb + pos, version, h * w, true, u) || (long long)pos + u.len > left) bad = true;
          else if (u.len > MAXU) fallback = true;
          else {
            if (fdPattern(sb[pos], version) != ((tx * 8 >> 3) & (version >= 5 ? 14 : 15))) bad = true;
            if (g == sbI) { myPos = pos; myBlk = (int)blk; }
            pos += u.len; blk++;
          }
        }
      }
      if (bad || fallback) break;
      // ---- decode my block's row r
      if (myPos >= 0) {
        const int ty = myBlk / a.nTx, tx = myBlk - ty * a.nTx;
        const int i0 = ty * 8, j0 = tx * 8;
        const int h = min(8, a.nRows - i0), w = min(8, a.nCols - j0);
        if (r < h) {
          FdUnit u;
          const uint8_t* p = sb + myPos;
          fdParse<T>(p, version, h * w, true, u);
          T out[8];
          if (u.mode == 2) {
#pragma unroll
            for (int k = 0; k < 8; k++) out[k] = (T)0;
          } else if (u.mode == 0) {
            const uint8_t* src = p + 1 + (size_t)(r * w) * sizeof(T);
#pragma unroll
            for (int k = 0; k < 8; k++) {
              T val = (T)0;
              if (k < w) { uint8_t* vb = (uint8_t*)&val;
#pragma unroll
                for (int bb = 0; bb < (int)sizeof(T); bb++) vb[bb] = src[k * sizeof(T) + bb]; }
              out[k] = val;
            }
          } else {
            const int dtUsed = offsetTypeFromCode(PixelTraits<T>::code, u.tc);
            const double offset = offsetFromBits(loadBytesLE(p + 1, u.osz), dtUsed);
            if (u.mode == 3) {
#pragma unroll
              for (int k = 0; k < 8; k++) out[k] = (T)offset;
            } else {
              const int nbv = u.lut ? u.nbIdx : u.nb;
              const uint32_t bit0 = (uint32_t)(r * w) * (uint32_t)nbv;
#pragma unroll
              for (int k = 0; k < 8; k++) {
                uint32_t q = 0;
                if (k < w && nbv > 0) q = fdExtract(p + u.pay, bit0 + (uint32_t)(k * nbv), nbv);
                if (u.lut) {
                  if (q > (uint32_t)u.nLut) { bad = true; q = 0; }
                  q = q == 0 ? 0u : fdExtract(p + u.lutPay, (q - 1) * (uint32_t)u.nb, u.nb);
                }
                const double z = __dadd_rn(offset, __dmul_rn((double)q, a.invScale));
                out[k] = fdCast<T>(z, a.zMax);
              }
            }
          }
          T* dst = data + (size_t)(i0 + r) * a.nCols + j0;
          if (w == 8 && vecOk) {
            if (sizeof(T) == 4) { uint32_t o[8]; memcpy(o, out, 32); ((uint4*)dst)[0] = make_uint4(o[0], o[1], o[2], o[3]); ((uint4*)dst)[1] = make_uint4(o[4], o[5], o[6], o[7]); }
            else if (sizeof(T) == 8) { uint32_t o[16]; memcpy(o, out, 64);
#pragma unroll
              for (int k = 0; k < 4; k++) ((uint4*)dst)[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]); }
            else if (sizeof(T) == 2) { uint32_t o[4]; memcpy(o, out, 16); ((uint4*)dst)[0] = make_uint4(o[0], o[1], o[2], o[3]); }
            else { uint32_t o[2]; memcpy(o, out, 8); ((uint2*)dst)[0] = make_uint2(o[0], o[1]); }
          } else {
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < w) dst[k] = out[k];
          }
        }
      }
      bad = __any_sync(FULL, bad);
    }
    // the serial parse continues exactly where the next sub-chunk's speculative chain started?
    if (!bad && !fallback) {
      const uint32_t gpos = (uint32_t)pos + (uint32_t)s * FD_SUB;
      if (blk < (uint32_t)nBlocks) {
        if (expectExit >= FD_DEAD - 1 || gpos != expectExit || blk != expectBlk) { fallback = true; }
      }
    }
    if (lane == 0 && bad) atomicOr(a.status, DECF_FALLBACK);      // the general decoder decides what is malformed
    if (lane == 0 && fallback) atomicOr(a.status, DECF_FALLBACK);
    bad = false; fallback = false;
    __syncwarp();
  }
  // the chain must cover all blocks: the last region (or whoever holds the tail) checks the block count
  if (tid == 0) {
    if (sBad) atomicOr(a.status, DECF_FALLBACK);
    if (reg == a.nReg - 1 && sTrue[2 * nLocal] != FD_DEAD && sTrue[2 * nLocal + 1] < (uint32_t)nBlocks) atomicOr(a.status, DECF_FALLBACK);
  }
}

template <class T> inline size_t fastDecodeSmemBytes(int subPerReg, int nReg) {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  constexpr int BUFB = ((FD_SUB + MAXU + 64 + 15) / 16) * 16;
  return (size_t)FD_WARPS * BUFB + (size_t)subPerReg * FD_ENT * sizeof(FdEntry) + (size_t)nReg * FD_ENT * sizeof(FdEntry) +
         (size_t)(subPerReg + 1) * 8 + (size_t)nReg * 4 + (size_t)FD_WARPS * 3 * FD_WK * 2 + (size_t)subPerReg + 16;
}

}  // namespace lerc
