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
  uint16_t* ckList;                      // [nSub][8] checkpoints of each sub-chunk's lowest surviving chain
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

// ---- header parse from a register window ---------------------------------------------------------
// The staged stream is read through aligned 32-bit shared-memory words; FdWin holds bytes p .. p+15 of it.
struct FdWin { unsigned long long lo, hi; };
__device__ __forceinline__ FdWin fdWindow(const uint8_t* __restrict__ sbase, int p) {      // sbase + p may have any alignment
  const uintptr_t ad = (uintptr_t)(sbase + p);
  const uint32_t* w = (const uint32_t*)(ad & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(ad & 3) * 8;
  const uint32_t a = w[0], b = w[1], c = w[2], d = w[3], e = w[4];
  FdWin x;
  x.lo = (unsigned long long)__funnelshift_r(a, b, sh) | ((unsigned long long)__funnelshift_r(b, c, sh) << 32);
  x.hi = (unsigned long long)__funnelshift_r(c, d, sh) | ((unsigned long long)__funnelshift_r(d, e, sh) << 32);
  return x;
}
__device__ __forceinline__ uint32_t fdByte(const FdWin& x, int i) { return (uint32_t)((i < 8 ? x.lo >> (8 * i) : x.hi >> (8 * (i - 8))) & 0xff); }

// Length of the unit starting at sbase[p] and its basic fields, from the window only (no LUT blocks: those take the
// byte-wise parser).  Returns 0 for a malformed header, -1 when the byte-wise parser is needed.
struct FdQuick { int mode, tc, osz, nb, n, pay, len; };
template <class T>
__device__ __forceinline__ int fdQuick(const FdWin& x, int version, int cells, bool exact, FdQuick& u) {
  constexpr int DT = PixelTraits<T>::code;
  const uint32_t flag = (uint32_t)x.lo & 0xff;
  u.mode = flag & 3; u.tc = flag >> 6; u.osz = 0; u.nb = 0; u.n = 0; u.pay = 1;
  if (version >= 5 && (flag & 4)) return 0;
  if (u.mode == 2) { u.len = 1; return 1; }
  if (u.mode == 0) { u.len = 1 + cells * (int)sizeof(T); return 1; }
  const int dtUsed = offsetTypeFromCode(DT, u.tc);
  if (dtUsed == DT_Undefined) return 0;
  u.osz = dtSize(dtUsed);
  if (u.mode == 3) { u.len = 1 + u.osz; return 1; }
  const uint32_t b = fdByte(x, 1 + u.osz), code = b >> 6;
  u.nb = b & 31;
  if ((b >> 5) & 1) return -1;                                       // LUT block
  if (code != 2) return code == 3 ? 0 : -1;                          // count field wider than one byte: byte-wise parser
  const uint32_t n = fdByte(x, 2 + u.osz);
  if (exact ? (n != (uint32_t)cells) : (n == 0 || n > (uint32_t)cells)) return 0;
  u.n = (int)n; u.pay = 3 + u.osz;
  u.len = u.pay + (int)packedBytes(n, u.nb);
  return 1;
}
// length only, any block kind
template <class T>
__device__ __forceinline__ int fdUnitLen(const uint8_t* __restrict__ sbase, int p, int version, int cells, bool exact, int& mode) {
  FdQuick q;
  const int rc = fdQuick<T>(fdWindow(sbase, p), version, cells, exact, q);
  mode = q.mode;
  if (rc > 0) return q.len;
  if (rc == 0) return 0;
  FdUnit u;
  if (!fdParse<T>(sbase + p, version, cells, exact, u)) return 0;
  return u.len;
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
  // the stream's last block may be a raw block of a partial tile, shorter than the 8x8 raw block the walkers assume
  const int tailRaw = 1 + (a.nRows - (a.nTy - 1) * 8) * (a.nCols - (a.nTx - 1) * 8) * (int)sizeof(T);
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
    // One hop of a speculative walker: length of the unit at pos, or 0.  A raw block that does not fit into the
    // rest of the stream is the shorter raw block of a partial tile at the very end (phase 4 checks exactly).
    auto hop = [&](int pos, int& np) -> int {
      int mode; int len = fdUnitLen<T>(sb, pos, version, 64, false, mode);
      np = fdPattern(sb[pos], version);
      if (len && (long long)pos + len > left) len = (mode == 0 && left - pos == tailRaw) ? (int)(left - pos) : 0;
      return len;
    };
    // ---- stage A: every position of the head window that parses as a header becomes a walker
    int nW = 0;
    const int headEnd = s == 0 ? 1 : min(MAXU, subEnd);                        // sub-chunk 0 starts with block 0 at position 0
    for (int base = 0; base < headEnd; base += 32) {
      const int p = base + lane;
      int len = 0, pat = 0;
      if (p < headEnd) len = hop(p, pat);
      const unsigned m = __ballot_sync(FULL, len > 0);
      if (len > 0) { const int i = nW + __popc(m & ((1u << lane) - 1)); wPos[i] = (uint16_t)(p + len); wEnt[i] = (uint16_t)p; wCnt[i] = (uint16_t)pat; }
      nW += __popc(m);
    }
    __syncwarp();
    // ---- stage B: hop with in-place stable compaction until at most 32 walkers are left (wrong candidates die fast)
    for (int pass = 0; nW > 32 && pass < 64; pass++) {
      int nNew = 0;
      for (int base = 0; base < nW; base += 32) {
        const int i = base + lane;
        bool alive = false; int pos = 0, ent = 0, pat = 0;
        if (i < nW) { pos = wPos[i]; ent = wEnt[i]; pat = wCnt[i]; alive = true; }
        if (alive && pos < subEnd) {
          int np; const int len = hop(pos, np);
          if (len > 0 && fdFollows(pat, np, version)) { pos += len; pat = np; } else alive = false;
        }
        __syncwarp();
        const unsigned ma = __ballot_sync(FULL, alive);
        if (alive) { const int j = nNew + __popc(ma & ((1u << lane) - 1)); wPos[j] = (uint16_t)pos; wEnt[j] = (uint16_t)ent; wCnt[j] = (uint16_t)pat; }
        nNew += __popc(ma);
        __syncwarp();
      }
      nW = nNew;
    }
    if (nW > 32) { nW = 32; if (lane == 0) atomicOr(a.status, DECF_FALLBACK | 16); }   // pathological: > 32 live chains
    // ---- stage C: the survivors restart from their entries in registers and record checkpoints:
    // ck[k] = first block start >= k * 512 on the chain (k = 0..7)
    const bool mine = lane < nW;
    const int ent = mine ? wEnt[lane] : 0;
    __syncwarp();
    uint16_t* ck = wPos;                                                        // [32][8], the walker arrays are free now
    int pos = ent, cnt = 0, pat = 0, seg = 0;
    bool walking = mine, done = false;
    while (__any_sync(FULL, walking)) {
      if (walking) {
        while (seg < 8 && pos >= seg * 512) { ck[lane * 8 + seg] = (uint16_t)pos; seg++; }
        if (pos >= subEnd) { walking = false; done = true; for (; seg < 8; seg++) ck[lane * 8 + seg] = (uint16_t)pos; }
        else {
          int np; const int len = hop(pos, np);
          if (len > 0 && (cnt == 0 || fdFollows(pat, np, version)) && cnt < 4096) { pos += len; cnt++; pat = np; }
          else walking = false;
        }
      }
    }
    // results: the done lanes, lowest entries first (lanes are sorted by entry)
    const unsigned md = __ballot_sync(FULL, done);
    const int rank = __popc(md & ((1u << lane) - 1)), nRes = __popc(md);
    for (int e = lane; e < FD_ENT; e += 32) sTab[(size_t)ls * FD_ENT + e].entry = FD_DEAD;
    __syncwarp();
    if (done && rank < FD_ENT) { FdEntry e; e.entry = (uint32_t)ent + (uint32_t)s * FD_SUB; e.exit = (uint32_t)pos + (uint32_t)s * FD_SUB; e.count = (uint32_t)cnt; sTab[(size_t)ls * FD_ENT + rank] = e; }
    if (lane == 0) sTabN[ls] = (uint8_t)min(nRes, FD_ENT);
    // checkpoints of the lowest surviving chain (later survivors are on the same chain once they have merged)
    if (md) {
      const int L = __ffs(md) - 1;
      if (lane < 8) a.ckList[(size_t)s * 8 + lane] = ck[L * 8 + lane];
    } else if (lane < 8) a.ckList[(size_t)s * 8 + lane] = 0xffff;
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
    if (lane == 0) { sRegEntry[0] = bad ? FD_DEAD : pos; sRegEntry[1] = blk; if (bad) sBad = 32; }
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
      if (!found) { sBad |= 64; pos = FD_DEAD; }
    }
  }
  __syncthreads();

  // ================= phase 4: decode ================================================================
  // A sub-chunk is cut at the checkpoints into up to 8 segments; the 4 lane groups of the warp take 4 segments
  // at a time: count the segment's blocks, prefix over the segments, then 8 lanes decode one block after the
  // other, lane r = block row r.
  T* data = (T*)a.data;
  const bool vecOk = ((a.nCols * (int)sizeof(T)) % 16 == 0) && (((uintptr_t)data & 15) == 0);
  const int g = lane >> 3, r = lane & 7;
  for (int ls = warp; ls < nLocal; ls += FD_WARPS) {
    const uint32_t pos0 = sTrue[2 * ls], blk0 = sTrue[2 * ls + 1];
    if (pos0 >= FD_DEAD - 1) continue;                                         // dead chain (reported through sBad) or past the end
    const int s = sub0 + ls;
    const uint32_t expectExit = sTrue[2 * ls + 2], expectBlk = sTrue[2 * ls + 3];
    const int d = stageSub(s);
    const uint8_t* sb = buf + d;
    const long long left = (long long)(a.streamLen - (unsigned long long)s * FD_SUB);
    const int subEnd = (int)min((long long)FD_SUB, left);
    const int p0 = (int)(pos0 - (uint32_t)s * FD_SUB);
    // where this sub-chunk's chain must end: the next sub-chunk's true entry; unknown (UNB) when the stream's last block lies in here
    constexpr int UNB = 0x7fffffff;
    const bool haveNext = expectExit < FD_DEAD - 1;
    const int exitRel = haveNext ? (int)(expectExit - (uint32_t)s * FD_SUB) : UNB;
    // segment boundaries: lane k (<= 8) holds bnd[k]; bnd[0] = p0, bnd[1..7] = checkpoints, bnd[8] = exit
    int myB;
    {
      int c = UNB;
      if (lane >= 1 && lane < 8) { const int v = (int)a.ckList[(size_t)s * 8 + lane]; c = v == 0xffff ? -1 : v; }
      if (lane == 0) c = p0;
      if (lane == 8) c = exitRel;
      const int prev = __shfl_up_sync(FULL, c, 1);
      bool okc = lane == 0 || lane > 8 || (c >= prev && c >= p0);     // non-decreasing, nothing before the true entry
      okc = __all_sync(FULL, okc);
      if (!okc && lane >= 1 && lane < 8) c = exitRel;                  // one segment: group 0 walks the whole sub-chunk
      myB = c;
    }
    bool fallback = false; unsigned why = 0;
    // ---- count the blocks of all 8 segments (lengths as in phase 1); when a checkpoint turns out not to lie on the
    // true chain the sub-chunk is redone as one segment
    int segS[2], segE[2], segN[2], segP[2];
    for (int attempt = 0; attempt < 2; attempt++) {
      bool mism = false;
#pragma unroll
      for (int round = 0; round < 2; round++) {
        const int k = round * 4 + g;
        const int segStart = __shfl_sync(FULL, myB, k), segEnd = __shfl_sync(FULL, myB, k + 1);
        int p = segStart, n = 0;
        while (p < segEnd && p < subEnd) {
          int mode; int len = fdUnitLen<T>(sb, p, version, 64, false, mode);
          if (len && (long long)p + len > left) len = (mode == 0 && left - p == tailRaw) ? (int)(left - p) : 0;
          if (!len || n >= 4096) { mism = true; break; }
          p += len; n++;
        }
        if (segEnd != UNB && segStart < segEnd && p != segEnd) mism = true;     // must land exactly on the next boundary
        segS[round] = segStart; segE[round] = segEnd; segN[round] = n; segP[round] = p;
      }
      mism = __any_sync(FULL, mism);
      if (!mism) break;
      if (attempt == 1) { fallback = true; why |= 256; break; }
      if (lane >= 1 && lane < 8) myB = exitRel;
    }
    uint32_t blkBase = blk0;
    for (int round = 0; round < 2 && !fallback; round++) {
      const int segStart = segS[round], segEnd = segE[round], n = segN[round], pCount = segP[round];
      (void)segEnd;
      int p;
      // ---- exclusive prefix of the counts over the 4 groups
      const int n0 = __shfl_sync(FULL, n, 0), n1 = __shfl_sync(FULL, n, 8), n2 = __shfl_sync(FULL, n, 16), n3 = __shfl_sync(FULL, n, 24);
      uint32_t b = blkBase + (g > 0 ? n0 : 0) + (g > 1 ? n1 : 0) + (g > 2 ? n2 : 0);
      blkBase += (uint32_t)(n0 + n1 + n2 + n3);
      // ---- decode my segment's blocks
      p = segStart;
      for (int i = 0; i < n && b < (uint32_t)nBlocks && !fallback; i++, b++) {
        const int ty = (int)b / a.nTx, tx = (int)b - ty * a.nTx;
        const int i0 = ty * 8, j0 = tx * 8;
        const int h = min(8, a.nRows - i0), w = min(8, a.nCols - j0), cells = h * w;
        const FdWin win = fdWindow(sb, p);
        FdQuick q;
        const int rc = fdQuick<T>(win, version, cells, true, q);
        if (rc == 0 || fdPattern((uint32_t)win.lo & 0xff, version) != (tx & (version >= 5 ? 14 : 15))) { fallback = true; why |= 512; break; }
        T out[8];
        int len;
        if (rc > 0) {
          len = q.len;
          if (q.mode == 2) {
#pragma unroll
            for (int kk = 0; kk < 8; kk++) out[kk] = (T)0;
          } else if (q.mode == 0) {
            const uint8_t* src = sb + p + 1 + (size_t)(r * w) * sizeof(T);
#pragma unroll
            for (int kk = 0; kk < 8; kk++) {
              T val = (T)0;
              if (kk < w && r < h) { uint8_t* vb = (uint8_t*)&val;
#pragma unroll
                for (int bb = 0; bb < (int)sizeof(T); bb++) vb[bb] = src[kk * sizeof(T) + bb]; }
              out[kk] = val;
            }
          } else {
            const int dtUsed = offsetTypeFromCode(PixelTraits<T>::code, q.tc);
            unsigned long long ob = win.lo >> 8;
            if (q.osz == 8) ob |= win.hi << 56;
            const double offset = offsetFromBits(q.osz == 8 ? ob : (ob & ((1ull << (8 * q.osz)) - 1)), dtUsed);
            if (q.mode == 3) {
#pragma unroll
              for (int kk = 0; kk < 8; kk++) out[kk] = (T)offset;
            } else {
              uint32_t qv[8];
              const int nb = q.nb;
              if (nb == 0) {
#pragma unroll
                for (int kk = 0; kk < 8; kk++) qv[kk] = 0;
              } else if (w == 8 && nb <= 16) {
                // the row is nb bytes at a byte boundary: 128-bit window, then split in halves / quarters / values
                const FdWin rw = fdWindow(sb, p + q.pay + r * nb);
                const int s4 = 4 * nb, s2 = 2 * nb;
                const unsigned long long m4 = s4 == 64 ? ~0ull : ((1ull << s4) - 1), m2 = (1ull << s2) - 1;
                const uint32_t m1 = (1u << nb) - 1;
                const unsigned long long h0 = rw.lo & m4;
                const unsigned long long h1 = (s4 == 64 ? rw.hi : ((rw.lo >> s4) | (rw.hi << (64 - s4)))) & m4;
                const unsigned long long q0 = h0 & m2, q1 = h0 >> s2, q2 = h1 & m2, q3 = h1 >> s2;
                qv[0] = (uint32_t)q0 & m1; qv[1] = (uint32_t)(q0 >> nb); qv[2] = (uint32_t)q1 & m1; qv[3] = (uint32_t)(q1 >> nb);
                qv[4] = (uint32_t)q2 & m1; qv[5] = (uint32_t)(q2 >> nb); qv[6] = (uint32_t)q3 & m1; qv[7] = (uint32_t)(q3 >> nb);
              } else {
                const uint32_t bit0 = (uint32_t)(r * w) * (uint32_t)nb;
#pragma unroll
                for (int kk = 0; kk < 8; kk++) qv[kk] = (kk < w && r < h) ? fdExtract(sb + p + q.pay, bit0 + (uint32_t)(kk * nb), nb) : 0u;
              }
#pragma unroll
              for (int kk = 0; kk < 8; kk++) out[kk] = fdCast<T>(__dadd_rn(offset, __dmul_rn((double)qv[kk], a.invScale)), a.zMax);
            }
          }
        } else {                                                             // LUT block or wide count field: byte-wise parser
          FdUnit u;
          if (!fdParse<T>(sb + p, version, cells, true, u) || u.len > MAXU) { fallback = true; why |= 8192; break; }
          len = u.len;
          const uint8_t* pp = sb + p;
          const int dtUsed = offsetTypeFromCode(PixelTraits<T>::code, u.tc);
          const double offset = offsetFromBits(loadBytesLE(pp + 1, u.osz), dtUsed);
          const int nbv = u.lut ? u.nbIdx : u.nb;
          const uint32_t bit0 = (uint32_t)(r * w) * (uint32_t)nbv;
          bool badLut = false;
#pragma unroll
          for (int kk = 0; kk < 8; kk++) {
            uint32_t qq = 0;
            if (kk < w && r < h && nbv > 0) qq = fdExtract(pp + u.pay, bit0 + (uint32_t)(kk * nbv), nbv);
            if (u.lut) {
              if (qq > (uint32_t)u.nLut) { badLut = true; qq = 0; }
              qq = qq == 0 ? 0u : fdExtract(pp + u.lutPay, (qq - 1) * (uint32_t)u.nb, u.nb);
            }
            out[kk] = fdCast<T>(__dadd_rn(offset, __dmul_rn((double)qq, a.invScale)), a.zMax);
          }
          if (badLut) { fallback = true; why |= 16384; }
        }
        if (len > MAXU) { fallback = true; why |= 32768; break; }
        if (r < h) {
          T* dst = data + (size_t)(i0 + r) * a.nCols + j0;
          if (w == 8 && vecOk) {
            if (sizeof(T) == 4) { uint32_t o[8]; memcpy(o, out, 32); ((uint4*)dst)[0] = make_uint4(o[0], o[1], o[2], o[3]); ((uint4*)dst)[1] = make_uint4(o[4], o[5], o[6], o[7]); }
            else if (sizeof(T) == 8) { uint32_t o[16]; memcpy(o, out, 64);
#pragma unroll
              for (int kk = 0; kk < 4; kk++) ((uint4*)dst)[kk] = make_uint4(o[4 * kk], o[4 * kk + 1], o[4 * kk + 2], o[4 * kk + 3]); }
            else if (sizeof(T) == 2) { uint32_t o[4]; memcpy(o, out, 16); ((uint4*)dst)[0] = make_uint4(o[0], o[1], o[2], o[3]); }
            else { uint32_t o[2]; memcpy(o, out, 8); ((uint2*)dst)[0] = make_uint2(o[0], o[1]); }
          } else {
#pragma unroll
            for (int kk = 0; kk < 8; kk++) if (kk < w) dst[kk] = out[kk];
          }
        }
        p += len;
      }
      // the exact walk (true block sizes) must end where the counting walk (phase-1 lengths) ended
      if (!fallback && b < (uint32_t)nBlocks && p != pCount) { fallback = true; why |= 1024; }
      if (__any_sync(FULL, fallback)) { fallback = true; break; }
    }
    // the serial parse continues in the next sub-chunk with the block index phase 3 assumed; or it ended in here
    if (!fallback) { if (haveNext ? (blkBase != expectBlk) : (blkBase < (uint32_t)nBlocks)) { fallback = true; why |= 2048; } }
    why = __reduce_or_sync(FULL, why);
    if (lane == 0 && fallback) atomicOr(a.status, DECF_FALLBACK | why);
    __syncwarp();
  }
  // the chain must cover all blocks: the last region (or whoever holds the tail) checks the block count
  if (tid == 0) {
    if (sBad) atomicOr(a.status, DECF_FALLBACK | sBad);
    if (reg == a.nReg - 1 && sTrue[2 * nLocal] != FD_DEAD && sTrue[2 * nLocal + 1] < (uint32_t)nBlocks) atomicOr(a.status, DECF_FALLBACK | 4096);
  }
}

template <class T> inline size_t fastDecodeSmemBytes(int subPerReg, int nReg) {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  constexpr int BUFB = ((FD_SUB + MAXU + 64 + 15) / 16) * 16;
  return (size_t)FD_WARPS * BUFB + (size_t)subPerReg * FD_ENT * sizeof(FdEntry) + (size_t)nReg * FD_ENT * sizeof(FdEntry) +
         (size_t)(subPerReg + 1) * 8 + (size_t)nReg * 4 + (size_t)FD_WARPS * 3 * FD_WK * 2 + (size_t)subPerReg + 16;
}

}  // namespace lerc
