// lerc_decode_fast.cuh -- parallel decoder of a Lerc2 micro-block stream for the common raster shape:
// every pixel valid, nDepth == 1, 8x8 micro-blocks (included by lerc_decode.cu).
//
// The stream has no index: a block's length is only known from its own header bytes (ReadTile, Lerc2.cpp:2025-2230;
// BitStuffer2::Decode, BitStuffer2.cpp:159-258), so the reference walks it serially.  Here the block boundaries are
// found speculatively, in four kernels on one stream:
//   k_dec_candidates  the stream is cut into 4 KB sub-chunks, one warp each.  Every byte position in the first MAXU
//                     bytes of a sub-chunk that parses as a block header is a candidate entry (32 positions per warp
//                     step, staged in shared memory); candidates hop from header to header and die on the first
//                     malformed header or broken integrity-bit sequence (Lerc2.cpp:2045) until at most FD_CAND are left.
//   k_dec_walk        one LANE per surviving candidate walks its chain to the end of the sub-chunk straight from
//                     L2 (the checksum kernel has just read the blob), recording the unit lengths; result per
//                     candidate: (entry, exit, #blocks).  Wrong candidates die or merge into the true chain.
//                     The sub-chunk maps of a region (one CTA) are composed in shared memory -> region map.
//   k_dec_resolve     one warp composes the region maps in stream order: true (position, block index) of every region.
//   k_dec_blocks      every CTA (region) resolves its sub-chunks' true entries; the recorded lengths of the matching candidate give every block's
//                     position, so the blocks decode in parallel: 8 lanes per micro-block, one lane per block row:
//                     unpack numBits-wide values with funnel shifts, z = offset + q * 2 maxZError in fp64 without
//                     contraction, min(z, zMax), cast, 128-bit stores.  Every block is re-parsed with its true size and
//                     must end exactly where the next one starts, every sub-chunk where the next one's chain starts:
//                     the result is the serial parse or an error flag -- never a silently different parse.
// Streams the speculation cannot handle (partial raw blocks inside the stream, units longer than MAXU, too many
// surviving candidates, corrupt data) raise DECF_FALLBACK and the host runs the general decoder (decodeBandT),
// which decides what is malformed exactly like the reference.
#pragma once

namespace lerc {

enum { DECF_FALLBACK = 8 };
constexpr int FD_SUB = 4096;          // sub-chunk bytes
constexpr int FD_CAND = 16;           // surviving entry candidates kept per sub-chunk / region
constexpr int FD_MAXHOP = 4096;       // blocks per sub-chunk a chain may hold (1-byte blocks fill a sub-chunk with 4096)
constexpr int FD_LENS = 512;          // bytes reserved per candidate for its unit lengths, run-length coded:
constexpr int FD_PAIRS = 252;         //   up to FD_PAIRS (length code, repeat) byte pairs, pair count as uint16 at byte FD_LENS - 2
constexpr int FD_BATCH = 512;         // block positions expanded at a time in k_dec_blocks / k_dec_offsets
constexpr uint32_t FD_DEAD = 0xffffffffu;
constexpr int FD_REG = 16;            // sub-chunks per region (one CTA of k_dec_walk / k_dec_blocks)

struct FdEntry { uint32_t entry, exit, count; };
struct FdCand { uint16_t entry, pos, cnt, pat; };

struct FastDecArgs {
  const uint8_t* stream; unsigned long long streamLen;
  int nRows, nCols, nTx, nTy, dt, version;
  double invScale, zMax;                 // 2 * maxZError ; header zMax
  void* data;
  int nSub, subPerReg, nReg;
  FdCand* cand; uint8_t* nCand;          // [nSub][FD_CAND], [nSub]
  uint8_t* lens;                         // [nSub][FD_CAND][FD_LENS] unit lengths of each candidate's chain
  FdEntry* subTab;                       // [nSub][FD_CAND]
  FdEntry* regTab;                       // [nReg][FD_CAND]
  uint32_t* regEntry;                    // [nReg + 1][2] true (position, block index) at the start of every region
  int* status;
  // repair launch of k_dec_walk (one CTA): the true chain enters region repairReg at stream position repairPos, which the
  // head-window speculation did not keep among the region's candidates (candidate flood); -1 = normal launch
  int repairReg; uint32_t repairPos;
  // experimental (LERC_B200_DEC=closure): head-window candidates only for the FIRST sub-chunk of every region; the entries of the other
  // sub-chunks are the real chain exits of their predecessors, walked by k_dec_walk's closure pass (one lane per distinct exit)
  int firstOnly;
};

// ---- unit header ---------------------------------------------------------------------------------
struct FdUnit { int mode, tc, osz, nb, lut, n, nLut, nbIdx, pay, lutPay, len; };

// Byte-wise parser of the unit whose first byte is p[0] (>= 16 readable bytes), any block kind.  cells = pixels of
// the block (every pixel is valid here), so raw blocks hold cells values and bit-stuffed blocks must hold exactly
// cells values (Lerc2.cpp:2148).  Returns false when the reference's ReadTile / BitStuffer2::Decode would fail.
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
// FdWin holds bytes p .. p+15 of the stream, assembled from aligned 32-bit words (shared or global memory).
struct FdWin { unsigned long long lo, hi; };
__device__ __forceinline__ FdWin fdWindow(const uint32_t* __restrict__ words, uint32_t byteOff) {
  const uint32_t* w = words + (byteOff >> 2);
  const uint32_t sh = (byteOff & 3) * 8;
  const uint32_t a = w[0], b = w[1], c = w[2], d = w[3], e = w[4];
  FdWin x;
  x.lo = (unsigned long long)__funnelshift_r(a, b, sh) | ((unsigned long long)__funnelshift_r(b, c, sh) << 32);
  x.hi = (unsigned long long)__funnelshift_r(c, d, sh) | ((unsigned long long)__funnelshift_r(d, e, sh) << 32);
  return x;
}
__device__ __forceinline__ uint32_t fdByte(const FdWin& x, int i) { return (uint32_t)((i < 8 ? x.lo >> (8 * i) : x.hi >> (8 * (i - 8))) & 0xff); }

// Basic fields and length of a unit from its window.  1 = parsed; 0 = malformed; -1 = LUT block or wide count
// field (the byte-wise parser is needed).
struct FdQuick { int mode, tc, osz, nb, n, pay, len; };
template <class T>
__device__ __forceinline__ int fdQuick(const FdWin& x, int version, int cells, bool exact, FdQuick& u) {
  constexpr int DT = PixelTraits<T>::code;
  const uint32_t flag = (uint32_t)x.lo & 0xff;
  u.mode = flag & 3; u.tc = flag >> 6; u.osz = 0; u.nb = 0; u.n = 0; u.pay = 1; u.len = 0;
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

// integrity bits of a block header (Lerc2.cpp:2045): (j0 >> 3) & 15, only bits 1..3 of it from version 5 on
__device__ __forceinline__ int fdPattern(unsigned flag, int version) { return (int)((flag >> 2) & (version >= 5 ? 14u : 15u)); }
// may a block with pattern b follow a block with pattern a in stream order?  (tx -> tx + 1, or a new block row)
__device__ __forceinline__ bool fdFollows(int a, int b, int version) {
  if (b == 0) return true;
  return version >= 5 ? (b == a || b == ((a + 2) & 14)) : (b == ((a + 1) & 15));
}

// bits [bit, bit + nb) of a little-endian bit stream (byte pointer into shared memory, any alignment)
__device__ __forceinline__ uint32_t fdExtract(const uint8_t* __restrict__ base, uint32_t bit, int nb) {
  const uint32_t byte = bit >> 3;
  const uint32_t* w = (const uint32_t*)((uintptr_t)(base + byte) & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(((uintptr_t)(base + byte) & 3) * 8 + (bit & 7));
  const unsigned long long x = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
  uint32_t v = (uint32_t)(x >> sh);
  if (sh + nb > 64) v |= w[2] << (64 - sh);
  return nb >= 32 ? v : (v & ((1u << nb) - 1));
}

template <class T> __device__ __forceinline__ T fdCast(double z, double zMax) { const double v = z < zMax ? z : zMax; return (T)v; }   // Lerc2.cpp:2160

// Speculative hop: length of the unit whose window is x (an 8x8 block is assumed), 0 = dead.  `rest` = stream bytes
// from this unit's first byte; a raw block that does not fit can only be the raw block of the partial tile that
// ends the stream (tailRaw bytes).  LUT / wide-count units take the byte-wise parser through `bytes` (16+ readable
// bytes) when it is not null, and end the chain otherwise.
template <class T>
__device__ __noinline__ int fdSlowLen(const uint8_t* bytes, int version) {
  FdUnit u;
  return fdParse<T>(bytes, version, 64, false, u) ? u.len : 0;
}
template <class T>
__device__ __forceinline__ int fdHopLen(const FdWin& x, const uint8_t* bytes, int version, long long rest, int tailRaw, int& pat) {
  FdQuick q;
  int rc = fdQuick<T>(x, version, 64, false, q);
  pat = fdPattern((uint32_t)x.lo & 0xff, version);
  int len = q.len;
  if (rc < 0) {
    len = bytes ? fdSlowLen<T>(bytes, version) : 0;
    rc = len > 0 ? 1 : 0;
  }
  if (rc <= 0) return 0;
  if ((long long)len > rest) return (q.mode == 0 && rest == (long long)tailRaw) ? tailRaw : 0;
  return len;
}

// One block row of a LUT block / a block with a wide count field (byte-wise parser).  Returns the unit length,
// 0 when malformed, -2 when a LUT index is out of range.
template <class T>
__device__ __noinline__ int fdDecodeSlowRow(const uint8_t* pp, int version, int cells, int r, int h, int w, double invScale, double zMax, T* out) {
  FdUnit u;
  if (!fdParse<T>(pp, version, cells, true, u)) return 0;
  const int dtUsed = offsetTypeFromCode(PixelTraits<T>::code, u.tc);
  const double offset = offsetFromBits(loadBytesLE(pp + 1, u.osz), dtUsed);
  const int nbv = u.lut ? u.nbIdx : u.nb;
  const uint32_t bit0 = (uint32_t)(r * w) * (uint32_t)nbv;
  bool badIdx = false;
  for (int kk = 0; kk < 8; kk++) {
    uint32_t qq = 0;
    if (kk < w && r < h && nbv > 0) qq = fdExtract(pp + u.pay, bit0 + (uint32_t)(kk * nbv), nbv);
    if (u.lut) {
      if (qq > (uint32_t)u.nLut) { badIdx = true; qq = 0; }
      qq = qq == 0 ? 0u : fdExtract(pp + u.lutPay, (qq - 1) * (uint32_t)u.nb, u.nb);
    }
    out[kk] = fdCast<T>(__dadd_rn(offset, __dmul_rn((double)qq, invScale)), zMax);
  }
  return badIdx ? -2 : u.len;
}

// ================= kernel 1: entry candidates of every sub-chunk ===================================
template <class T>
__global__ void __launch_bounds__(256) k_dec_candidates(FastDecArgs a) {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  constexpr int STG = ((3 * MAXU + 48 + 15) / 16) * 16;             // staged bytes per warp (head window + two hops of look-ahead)
  constexpr int WK = ((MAXU + 31) / 32) * 32;                       // walker slots
  __shared__ __align__(16) uint8_t sStage[8][STG + 16];
  __shared__ uint16_t sW[8][4][WK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 8 + warp;
  if (s >= a.nSub) return;
  if (a.firstOnly && (s % a.subPerReg) != 0) { if (lane == 0) a.nCand[s] = 0; return; }
  const int version = a.version;
  const int tailRaw = 1 + (a.nRows - (a.nTy - 1) * 8) * (a.nCols - (a.nTx - 1) * 8) * (int)sizeof(T);
  const unsigned long long start = (unsigned long long)s * FD_SUB;
  const long long left = (long long)(a.streamLen - start);
  const int subEnd = (int)min((long long)FD_SUB, left);
  // stage: buf[d + i] = stream[start + i], d = global misalignment
  uint8_t* buf = sStage[warp];
  const uint8_t* g = a.stream + start;
  const int d = (int)((uintptr_t)g & 15);
  {
    const uint8_t* g0 = g - d;
    const long long avail = left + d;
    for (int i = lane; i < (STG + 16) / 16; i += 32) {
      uint4 x = make_uint4(0, 0, 0, 0);
      if ((long long)i * 16 < avail) x = __ldg((const uint4*)g0 + i);
      ((uint4*)buf)[i] = x;
    }
  }
  __syncwarp();
  const uint32_t* words = (const uint32_t*)buf;
  const int readable = STG - 24;                                     // a window at byte p needs p + d + 20 <= STG + 16
  uint16_t* wPos = sW[warp][0], *wEnt = sW[warp][1], *wCnt = sW[warp][2], *wPat = sW[warp][3];
  // ---- every position of the head window that parses as a header becomes a walker (already one hop ahead)
  int nW = 0;
  const int headEnd = s == 0 ? 1 : min(MAXU, subEnd);               // sub-chunk 0 starts with block 0 at position 0
  for (int base = 0; base < headEnd; base += 32) {
    const int p = base + lane;
    int len = 0, pat = 0;
    if (p < headEnd) len = fdHopLen<T>(fdWindow(words, (uint32_t)(d + p)), buf + d + p, version, left - p, tailRaw, pat);
    const unsigned m = __ballot_sync(FULL, len > 0);
    if (len > 0) { const int i = nW + __popc(m & ((1u << lane) - 1)); wPos[i] = (uint16_t)(p + len); wEnt[i] = (uint16_t)p; wCnt[i] = 1; wPat[i] = (uint16_t)pat; }
    nW += __popc(m);
  }
  __syncwarp();
  // ---- hop with in-place stable compaction until at most FD_CAND walkers are left (wrong candidates die fast)
  for (int pass = 0; nW > FD_CAND - 1 && pass < 8; pass++) {
    int nNew = 0; bool moved = false;
    for (int base = 0; base < nW; base += 32) {
      const int i = base + lane;
      bool alive = false; int pos = 0, ent = 0, cnt = 0, pat = 0;
      if (i < nW) { pos = wPos[i]; ent = wEnt[i]; cnt = wCnt[i]; pat = wPat[i]; alive = true; }
      if (alive && pos < subEnd && pos < readable) {
        int np; const int len = fdHopLen<T>(fdWindow(words, (uint32_t)(d + pos)), buf + d + pos, version, left - pos, tailRaw, np);
        if (len > 0 && fdFollows(pat, np, version)) { pos += len; pat = np; cnt++; moved = true; } else alive = false;
      }
      __syncwarp();
      const unsigned ma = __ballot_sync(FULL, alive);
      if (alive) { const int j = nNew + __popc(ma & ((1u << lane) - 1)); wPos[j] = (uint16_t)pos; wEnt[j] = (uint16_t)ent; wCnt[j] = (uint16_t)cnt; wPat[j] = (uint16_t)pat; }
      nNew += __popc(ma);
      __syncwarp();
    }
    nW = nNew;
    if (!__any_sync(FULL, moved)) break;
  }
  // more than FD_CAND left: keep the lowest entries (the true entry is the lowest position on the true chain)
  const int nKeep = min(nW, FD_CAND - 1);                           // one slot stays free for the closure pass of k_dec_walk
  if (lane < nKeep) { FdCand c; c.entry = wEnt[lane]; c.pos = wPos[lane]; c.cnt = wCnt[lane]; c.pat = wPat[lane]; a.cand[(size_t)s * FD_CAND + lane] = c; }
  if (lane == 0) a.nCand[s] = (uint8_t)nKeep;
}

// ================= kernel 2: one lane per candidate walks to the end of its sub-chunk ================
// CTA = region (subPerReg consecutive sub-chunks); afterwards warp 0 composes the region's sub-chunk maps.
// Walks the chain that starts at byte `entry` of sub-chunk s to the end of the sub-chunk (bytes of the region staged in
// shared memory: stagedBytes[base + p] = stream[s * FD_SUB + p]), recording the unit lengths as (code, repeat) pairs: code
// 255 stands for the raw 8x8 block (the only unit that can be longer than 254 bytes); flat regions (1..5-byte blocks,
// thousands per sub-chunk) collapse into a few pairs.  Returns false when the chain dies.
template <class T>
__device__ __forceinline__ bool fdWalkChain(const FastDecArgs& a, const uint8_t* __restrict__ stagedBytes, int base, int s, int entry,
                                            int version, int tailRaw, uint8_t* __restrict__ lens, FdEntry& e) {
  const uint32_t* words = (const uint32_t*)stagedBytes;
  const unsigned long long start = (unsigned long long)s * FD_SUB;
  const long long left = (long long)(a.streamLen - start);
  const int subEnd = (int)min((long long)FD_SUB, left);
  int pos = entry, cnt = 0, pat = 0;
  bool ok = true;
  unsigned long long acc = 0;
  int nPairs = 0, curCode = -1, curRun = 0;
  auto flushPair = [&]() {
    acc |= (unsigned long long)((unsigned)curCode | ((unsigned)curRun << 8)) << (16 * (nPairs & 3));
    if ((nPairs & 3) == 3) { *(unsigned long long*)(lens + (nPairs & ~3) * 2) = acc; acc = 0; }
    nPairs++;
  };
  while (pos < subEnd) {
    const FdWin x = fdWindow(words, (uint32_t)(base + pos));
    int np;
    const int len = fdHopLen<T>(x, (left - pos >= 24) ? stagedBytes + base + pos : nullptr, version, left - pos, tailRaw, np);   // byte-wise parser (LUT blocks) reads the staged bytes
    const int code = len == 1 + 64 * (int)sizeof(T) ? 255 : len;
    if (len <= 0 || (code != 255 && len >= 255) || cnt >= FD_MAXHOP || (cnt > 0 && !fdFollows(pat, np, version))) { ok = false; break; }
    if (code == curCode && curRun < 255) curRun++;
    else {
      if (curCode >= 0) { if (nPairs >= FD_PAIRS) { ok = false; break; } flushPair(); }
      curCode = code; curRun = 1;
    }
    pos += len; cnt++; pat = np;
  }
  if (ok && curCode >= 0) { if (nPairs >= FD_PAIRS) ok = false; else flushPair(); }
  if (!ok) return false;
  if (nPairs & 3) *(unsigned long long*)(lens + (nPairs & ~3) * 2) = acc;
  *(uint16_t*)(lens + FD_LENS - 2) = (uint16_t)nPairs;
  e.entry = (uint32_t)entry + (uint32_t)s * FD_SUB; e.exit = (uint32_t)pos + (uint32_t)s * FD_SUB; e.count = (uint32_t)cnt;
  return true;
}

template <class T>
__global__ void __launch_bounds__(256) k_dec_walk(FastDecArgs a) {
  __shared__ FdEntry sTab[FD_REG * FD_CAND];                        // [subPerReg <= FD_REG][FD_CAND]
  extern __shared__ __align__(16) uint8_t sReg[];                    // the region's bytes: FD_REG * FD_SUB + look-ahead
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  constexpr int AHEAD = ((MAXU + 64 + 15) / 16) * 16;
  const int reg = a.repairReg >= 0 ? a.repairReg : (int)blockIdx.x, tid = threadIdx.x;
  const int sub0 = reg * a.subPerReg, nLocal = max(0, min(a.nSub, sub0 + a.subPerReg) - sub0);
  const int version = a.version;
  const int tailRaw = 1 + (a.nRows - (a.nTy - 1) * 8) * (a.nCols - (a.nTx - 1) * 8) * (int)sizeof(T);
  // ---- stage the region (the walks are dependent chains of header reads: shared-memory latency instead of L2 latency)
  const unsigned long long regStart = (unsigned long long)sub0 * FD_SUB;
  const uint8_t* g = a.stream + regStart;
  const int d = (int)((uintptr_t)g & 15);
  {
    const uint8_t* g0 = g - d;
    const long long avail = (long long)(a.streamLen - regStart) + d;
    const int nChunks = (nLocal * FD_SUB + AHEAD + 16) / 16;
    for (int i = tid; i < nChunks; i += blockDim.x) {
      uint4 x = make_uint4(0, 0, 0, 0);
      if ((long long)i * 16 < avail) x = __ldg((const uint4*)g0 + i);
      ((uint4*)sReg)[i] = x;
    }
  }
  __syncthreads();
  const uint32_t* words = (const uint32_t*)sReg;
  (void)words;
  if (a.repairReg >= 0) {
    // repair launch: the tables of this region exist; the true entry of the region is walked into a free slot of the first
    // sub-chunk, then closure and composition below run again
    for (int it = tid; it < nLocal * FD_CAND; it += blockDim.x) sTab[it] = a.subTab[(size_t)sub0 * FD_CAND + it];
    __syncthreads();
    if (tid == 0 && nLocal > 0 && a.repairPos >= (uint32_t)sub0 * FD_SUB && a.repairPos < (uint32_t)(sub0 + 1) * FD_SUB) {
      bool have = false; int slot = -1;
      for (int e2 = 0; e2 < FD_CAND; e2++) { const uint32_t en = sTab[e2].entry; have |= en == a.repairPos; if (en == FD_DEAD && slot < 0) slot = e2; }
      FdEntry w;
      if (!have && slot >= 0 &&
          fdWalkChain<T>(a, sReg, d, sub0, (int)(a.repairPos - (uint32_t)sub0 * FD_SUB), version, tailRaw, a.lens + ((size_t)sub0 * FD_CAND + slot) * FD_LENS, w)) {
        sTab[slot] = w; a.subTab[(size_t)sub0 * FD_CAND + slot] = w;
      }
    }
  } else
  for (int it = tid; it < nLocal * FD_CAND; it += blockDim.x) {
    const int ls = it / FD_CAND, j = it - ls * FD_CAND, s = sub0 + ls;
    FdEntry e; e.entry = FD_DEAD; e.exit = 0; e.count = 0;
    if (j < (int)a.nCand[s]) {
      const FdCand c = a.cand[(size_t)s * FD_CAND + j];
      // restart from the entry so that every unit length of the chain is recorded
      FdEntry w;
      if (fdWalkChain<T>(a, sReg, d + ls * FD_SUB, s, (int)c.entry, version, tailRaw, a.lens + ((size_t)s * FD_CAND + j) * FD_LENS, w)) e = w;
    }
    sTab[it] = e;
    a.subTab[(size_t)sub0 * FD_CAND + it] = e;
  }
  __syncthreads();
  // ---- closure inside the region: every chain exit of sub-chunk ls - 1 must be an entry of sub-chunk ls.  The head-window
  // speculation usually provides it; where it does not (e.g. when more wrong candidates than FD_CAND merged into the true chain
  // and the true entry was not among the kept ones) the missing chain is walked now.  Sequential over the sub-chunks, lanes =
  // the 16 exits of the previous one.
  // quick parallel check first (one warp per pair of neighbouring sub-chunks); the sequential pass runs only if something is missing
  __shared__ int sMissing;
  if (tid == 0) sMissing = (a.repairReg >= 0 || a.firstOnly) ? 1 : 0;
  __syncthreads();
  for (int ls = 1 + (tid >> 5); ls < nLocal; ls += blockDim.x >> 5) {
    const int lane = tid & 31, s = sub0 + ls;
    uint32_t x = FD_DEAD;
    if (lane < FD_CAND) { const FdEntry t = sTab[(ls - 1) * FD_CAND + lane]; if (t.entry != FD_DEAD) x = t.exit; }
    const bool inSub = x != FD_DEAD && (unsigned long long)x < a.streamLen && x >= (uint32_t)s * FD_SUB && x < (uint32_t)(s + 1) * FD_SUB;
    bool present = !inSub;
    if (inSub) for (int e2 = 0; e2 < FD_CAND; e2++) present |= sTab[ls * FD_CAND + e2].entry == x;
    if (__any_sync(FULL, !present) && lane == 0) sMissing = 1;
  }
  __syncthreads();
  if (tid < 32 && sMissing) {
    const int lane = tid;
    for (int ls = 1; ls < nLocal; ls++) {
      const int s = sub0 + ls;
      uint32_t x = FD_DEAD;
      if (lane < FD_CAND) { const FdEntry t = sTab[(ls - 1) * FD_CAND + lane]; if (t.entry != FD_DEAD) x = t.exit; }
      const bool inSub = x != FD_DEAD && (unsigned long long)x < a.streamLen && x >= (uint32_t)s * FD_SUB && x < (uint32_t)(s + 1) * FD_SUB;
      bool present = !inSub;
      if (inSub) for (int e2 = 0; e2 < FD_CAND; e2++) present |= sTab[ls * FD_CAND + e2].entry == x;
      unsigned miss = __ballot_sync(FULL, !present);
      if (a.firstOnly) {
        // every DISTINCT missing exit is walked by its own lane into its own free slot of this sub-chunk
        bool leader = !present;
        for (int o = 0; o < FD_CAND; o++) { const uint32_t xo = __shfl_sync(FULL, x, o); const bool mo = (miss >> o) & 1; if (mo && o < lane && xo == x) leader = false; }
        const unsigned leaders = __ballot_sync(FULL, leader);
        const unsigned freeSlots = __ballot_sync(FULL, lane < FD_CAND && sTab[ls * FD_CAND + lane].entry == FD_DEAD);
        if (leader) {
          int k = __popc(leaders & ((1u << lane) - 1));
          unsigned f = freeSlots;
          while (k > 0 && f) { f &= f - 1; k--; }
          if (f) {
            const int slot = __ffs(f) - 1;
            FdEntry w;
            if (fdWalkChain<T>(a, sReg, d + ls * FD_SUB, s, (int)(x - (uint32_t)s * FD_SUB), version, tailRaw, a.lens + ((size_t)s * FD_CAND + slot) * FD_LENS, w)) {
              sTab[ls * FD_CAND + slot] = w; a.subTab[(size_t)s * FD_CAND + slot] = w;
            }
          }
        }
        __syncwarp();
        miss = 0;
      }
      while (miss) {
        const int src = __ffs(miss) - 1;
        miss &= miss - 1;
        const uint32_t xs = __shfl_sync(FULL, x, src);
        if (lane == 0) {
          bool have = false; int slot = -1;
          for (int e2 = 0; e2 < FD_CAND; e2++) { const uint32_t en = sTab[ls * FD_CAND + e2].entry; have |= en == xs; if (en == FD_DEAD && slot < 0) slot = e2; }
          if (!have && slot >= 0) {
            FdEntry w;
            if (fdWalkChain<T>(a, sReg, d + ls * FD_SUB, s, (int)(xs - (uint32_t)s * FD_SUB), version, tailRaw, a.lens + ((size_t)s * FD_CAND + slot) * FD_LENS, w)) {
              sTab[ls * FD_CAND + slot] = w; a.subTab[(size_t)s * FD_CAND + slot] = w;
            }
          }
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
  // ---- compose the sub-chunk maps of this region: lane j follows the chain that starts at entry j of the first sub-chunk
  if (tid < 32) {
    const int lane = tid;
    FdEntry cur; cur.entry = FD_DEAD; cur.exit = 0; cur.count = 0;
    if (lane < FD_CAND && nLocal > 0) cur = sTab[lane];
    for (int ls = 1; ls < nLocal; ls++) {
      if (cur.entry != FD_DEAD && (unsigned long long)cur.exit < a.streamLen) {      // a chain that reached the end of the stream is complete
        bool found = false;
        for (int e = 0; e < FD_CAND; e++) {
          const FdEntry t = sTab[ls * FD_CAND + e];
          if (t.entry != FD_DEAD && t.entry == cur.exit) { cur.exit = t.exit; cur.count += t.count; found = true; break; }
        }
        if (!found) cur.entry = FD_DEAD;
      }
    }
    if (lane < FD_CAND) a.regTab[(size_t)reg * FD_CAND + lane] = cur;
  }
}

// ---- block positions from a candidate's run-length coded unit lengths -----------------------------------
// fdLoadPairs gives every (length, repeat) pair its first block index and first byte position (per-warp table in
// shared memory: [0..255] length, [256..511] first index, [512..767] first position, [768..1023] repeat).
constexpr int FD_PTAB = 1024;
template <int MAXU>
__device__ __forceinline__ int fdLoadPairs(const uint8_t* __restrict__ lens, uint32_t firstPos, int lane, uint16_t* __restrict__ tab) {
  const int nPairs = (int)*(const uint16_t*)(lens + FD_LENS - 2);
  unsigned long long raw[2] = {0ull, 0ull};
  if (lane * 8 < nPairs) { raw[0] = *(const unsigned long long*)(lens + lane * 16); raw[1] = *(const unsigned long long*)(lens + lane * 16 + 8); }
  uint32_t nb = 0, by = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint32_t pr = (uint32_t)(raw[k >> 2] >> (16 * (k & 3))) & 0xffffu;
    const uint32_t code = pr & 0xff, run = (lane * 8 + k < nPairs) ? (pr >> 8) : 0u;
    nb += run; by += run * (code == 255 ? (uint32_t)MAXU : code);
  }
  uint32_t inb = nb, iby = by;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    const uint32_t o1 = __shfl_up_sync(FULL, inb, m), o2 = __shfl_up_sync(FULL, iby, m);
    if (lane >= m) { inb += o1; iby += o2; }
  }
  uint32_t bi = inb - nb, pi = firstPos + iby - by;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint32_t pr = (uint32_t)(raw[k >> 2] >> (16 * (k & 3))) & 0xffffu;
    const uint32_t code = pr & 0xff, run = (lane * 8 + k < nPairs) ? (pr >> 8) : 0u;
    const uint32_t len = code == 255 ? (uint32_t)MAXU : code;
    const int j = lane * 8 + k;
    if (j < nPairs) { tab[j] = (uint16_t)len; tab[256 + j] = (uint16_t)bi; tab[512 + j] = (uint16_t)pi; tab[768 + j] = (uint16_t)run; }
    bi += run; pi += run * len;
  }
  __syncwarp();
  return nPairs;
}
// sPos[i - batchBase] = start of block i for batchBase <= i <= min(cnt, batchBase + FD_BATCH) (index cnt = the chain's exit)
__device__ __forceinline__ void fdFillBatch(const uint16_t* __restrict__ tab, int nPairs, int batchBase, int cnt, uint32_t exitRel, int lane, uint16_t* __restrict__ sPos) {
  const int hiIdx = min(cnt, batchBase + FD_BATCH);                  // last index wanted (inclusive)
  for (int j = lane; j < nPairs; j += 32) {
    const int run = tab[768 + j];
    if (!run) continue;
    const int idx0 = tab[256 + j], len = tab[j], pos0 = tab[512 + j];
    const int i0 = max(idx0, batchBase), i1 = min(idx0 + run, hiIdx + 1);
    for (int i = i0; i < i1; i++) sPos[i - batchBase] = (uint16_t)(pos0 + (i - idx0) * len);
  }
  if (lane == 0 && cnt >= batchBase && cnt <= hiIdx) sPos[cnt - batchBase] = (uint16_t)exitRel;
  __syncwarp();
}

// ================= kernel 3: true entry of every region ==============================================
// Region r's map sends each of its FD_CAND entry positions to (exit position, #blocks); the true chain starts at
// (position 0, block 0) in region 0.  One CTA of 32 warps; warp w owns the consecutive regions [w*G, (w+1)*G):
//   A  lanes 16..31 follow the 16 chains that start at the entries of the warp's first region through its G regions
//      (lanes 0..15 broadcast the next region's entries, every chain lane looks for its position) -> chunk map in shared memory
//   B  one thread composes the 32 chunk maps serially                        -> true entry of every chunk
//   C  every warp walks its regions again with the single true chain         -> regEntry[r] for every region
static_assert(FD_CAND == 16, "k_dec_resolve pairs 16 entry lanes with 16 chain lanes");
template <bool SMEM>
__global__ void __launch_bounds__(1024) k_dec_resolve(FastDecArgs a, int nBlocks) {
  // SMEM (experimental, LERC_B200_DEC_RESOLVE=smem): the region tables are copied to shared memory first, so that the ~56 dependent
  // steps below wait for shared memory instead of L2 (the kernel is one CTA of pure latency)
  extern __shared__ __align__(16) uint8_t sRegTabRaw[];
  const FdEntry* tab = a.regTab;
  if (SMEM) {
    FdEntry* sTabAll = (FdEntry*)sRegTabRaw;
    for (int i = threadIdx.x; i < a.nReg * FD_CAND; i += blockDim.x) sTabAll[i] = a.regTab[i];
    __syncthreads();
    tab = sTabAll;
  }
  __shared__ uint32_t sEnt[32][FD_CAND], sExit[32][FD_CAND], sCnt[32][FD_CAND];
  __shared__ uint32_t sChunkPos[33], sChunkBlk[33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (a.nReg + 31) / 32;
  const int r0 = warp * G, r1 = min(a.nReg, r0 + G);
  // ---- A
  {
    const int L = lane & 15;
    uint32_t ent = FD_DEAD, pos = FD_DEAD, cnt = 0;
    if (r0 < r1) { const FdEntry t = tab[(size_t)r0 * FD_CAND + L]; ent = t.entry; pos = t.exit; cnt = t.count; }
    bool alive = ent != FD_DEAD;
    FdEntry nx; nx.entry = FD_DEAD; nx.exit = 0; nx.count = 0;
    if (r0 + 1 < r1) nx = tab[(size_t)(r0 + 1) * FD_CAND + L];
    for (int r = r0 + 1; r < r1; r++) {
      const FdEntry t = nx;                                          // lanes 0..15 and 16..31 hold the same 16 entries
      if (r + 1 < r1) nx = tab[(size_t)(r + 1) * FD_CAND + L];  // independent of the chains: overlaps the match
      const bool chain = lane >= 16;
      const bool active = chain && alive && (unsigned long long)pos < a.streamLen;   // a chain that reached the end of the stream is complete
      int hit = -1;
#pragma unroll
      for (int j = 0; j < FD_CAND; j++) {                           // 16 independent shuffles (entry lanes broadcast their entry)
        const uint32_t e = __shfl_sync(FULL, t.entry, j);
        if (e != FD_DEAD && e == pos) hit = j;
      }
      const bool m = hit >= 0;
      const int src = m ? hit : 0;
      const uint32_t nx2 = __shfl_sync(FULL, t.exit, src), nc = __shfl_sync(FULL, t.count, src);
      if (active) { if (m) { pos = nx2; cnt += nc; } else alive = false; }
    }
    if (lane >= 16) { sEnt[warp][L] = alive ? ent : FD_DEAD; sExit[warp][L] = pos; sCnt[warp][L] = cnt; }
  }
  __syncthreads();
  // ---- B (warp 0: lanes 0..15 offer a chunk's entries, one ballot per chunk)
  if (warp == 0) {
    uint32_t pos = 0, blk = 0; bool dead = false;
    for (int w = 0; w < 32; w++) {
      if (lane == 0) { sChunkPos[w] = dead ? FD_DEAD : pos; sChunkBlk[w] = blk; }
      if (w * G >= a.nReg || dead || blk >= (uint32_t)nBlocks) continue;
      const uint32_t e = lane < FD_CAND ? sEnt[w][lane] : FD_DEAD;
      const unsigned m = __ballot_sync(FULL, e != FD_DEAD && e == pos);
      if (!m) { dead = true; continue; }
      const int src = __ffs(m) - 1;
      pos = sExit[w][src]; blk += sCnt[w][src];
    }
    if (lane == 0) {
      sChunkPos[32] = dead ? FD_DEAD : pos; sChunkBlk[32] = blk;
      a.regEntry[2 * a.nReg] = sChunkPos[32]; a.regEntry[2 * a.nReg + 1] = blk;
      if (dead) atomicOr(a.status, DECF_FALLBACK | 32);
      else if (blk < (uint32_t)nBlocks) atomicOr(a.status, DECF_FALLBACK | 4096);      // the chain must cover all blocks
    }
  }
  __syncthreads();
  // ---- C
  {
    uint32_t pos = sChunkPos[warp], blk = sChunkBlk[warp];
    bool dead = pos == FD_DEAD;
    FdEntry nx; nx.entry = FD_DEAD; nx.exit = 0; nx.count = 0;
    if (lane < FD_CAND && r0 < r1) nx = tab[(size_t)r0 * FD_CAND + lane];
    for (int r = r0; r < r1; r++) {
      const FdEntry t = nx;
      nx.entry = FD_DEAD;
      if (lane < FD_CAND && r + 1 < r1) nx = tab[(size_t)(r + 1) * FD_CAND + lane];   // independent of the chain: overlaps the lookup
      if (lane == 0) { a.regEntry[2 * r] = dead ? FD_DEAD : (blk >= (uint32_t)nBlocks ? FD_DEAD - 1 : pos); a.regEntry[2 * r + 1] = blk; }
      if (dead || blk >= (uint32_t)nBlocks) continue;
      const unsigned m = __ballot_sync(FULL, t.entry == pos && t.entry != FD_DEAD);
      if (!m) { dead = true; continue; }
      const int src = __ffs(m) - 1;
      pos = __shfl_sync(FULL, t.exit, src); blk += __shfl_sync(FULL, t.count, src);
    }
  }
}

// One row (lane r of the block's 8 lanes) of the block whose unit starts at byte p of the staged stream (`words` = the staging
// buffer as aligned words, sb = words + d bytes = stream byte 0 of the stage).  The unit is parsed with the block's true size.
// Returns the unit length; on a malformed / unsupported unit returns 0 and sets `why` (the caller falls back to the general decoder).
template <class T>
__device__ __forceinline__ int fdDecodeBlockRow(const uint32_t* __restrict__ words, const uint8_t* __restrict__ sb, int d, int p, int version, int patExpect,
                                                int cells, int h, int w, int r, double invScale, double zMax, T (&out)[8], unsigned& why) {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  const FdWin win = fdWindow(words, (uint32_t)(d + p));
  FdQuick q;
  const int rc = fdQuick<T>(win, version, cells, true, q);
  int len = 0;
  if (rc == 0 || fdPattern((uint32_t)win.lo & 0xff, version) != patExpect) { why |= 512; }
  else if (rc > 0 && q.len > MAXU) { why |= 32768; }                  // longer than the staged look-ahead: general decoder
  else if (rc > 0) {
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
          const FdWin rw = fdWindow(words, (uint32_t)(d + p + q.pay + r * nb));
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
        for (int kk = 0; kk < 8; kk++) out[kk] = fdCast<T>(__dadd_rn(offset, __dmul_rn((double)qv[kk], invScale)), zMax);
      }
    }
  } else {                                                     // LUT block or wide count field: byte-wise parser, out of line
    len = fdDecodeSlowRow<T>(sb + p, version, cells, r, h, w, invScale, zMax, out);
    if (len <= 0 || len > MAXU) { why |= len == -2 ? 16384 : 8192; len = 0; }
  }
  return len;
}

// 8 decoded pixels of one block row -> the raster (128-bit stores when the row is whole and aligned)
template <class T>
__device__ __forceinline__ void fdStoreRow(T* __restrict__ dst, const T (&out)[8], int w, bool vecOk) {
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

// True (position, block index, candidate slot) at the start of every sub-chunk of region `reg`, by one warp: the chain is followed
// sub-chunk by sub-chunk (a dependent walk of <= FD_REG steps), the FD_CAND candidates of a step are compared by 16 lanes at once.
__device__ __forceinline__ void fdTrueEntries(const FastDecArgs& a, int reg, int nLocal, int nBlocks, const FdEntry* __restrict__ sTab,
                                              uint32_t* __restrict__ sTrue, int* __restrict__ sWhy, int lane) {
  uint32_t pos = a.regEntry[2 * reg], blk = a.regEntry[2 * reg + 1];
  for (int ls = 0; ls <= nLocal; ls++) {
    if (lane == 0) { sTrue[3 * ls] = pos; sTrue[3 * ls + 1] = blk; sTrue[3 * ls + 2] = 0; }
    if (ls == nLocal) break;
    if (pos == FD_DEAD) { if (lane == 0) for (int k = ls; k <= nLocal; k++) { sTrue[3 * k] = FD_DEAD; sTrue[3 * k + 1] = blk; } break; }
    if (pos == FD_DEAD - 1 || blk >= (uint32_t)nBlocks) {                        // past the last block
      if (lane == 0) for (int k = ls; k <= nLocal; k++) { sTrue[3 * k] = FD_DEAD - 1; sTrue[3 * k + 1] = blk; }
      break;
    }
    FdEntry t; t.entry = FD_DEAD; t.exit = 0; t.count = 0;
    if (lane < FD_CAND) t = sTab[ls * FD_CAND + lane];
    const unsigned m = __ballot_sync(FULL, t.entry != FD_DEAD && t.entry == pos);
    if (!m) { if (lane == 0) *sWhy |= 64; pos = FD_DEAD; continue; }
    const int src = __ffs(m) - 1;                                                 // the lowest matching slot, like the serial search
    if (lane == 0) sTrue[3 * ls + 2] = (uint32_t)src;
    pos = __shfl_sync(FULL, t.exit, src); blk += __shfl_sync(FULL, t.count, src);
  }
  __syncwarp();
}

constexpr int FD_DWARPS = 8;
// ================= kernel 4b: block offsets only (masked rasters: the general block decoder does the pixels) =========
// Same resolution as k_dec_blocks; instead of decoding, every block's stream offset is written to blockOff[] for
// k_tiles_decode.  Bit-stuffed units describe their own length; a raw unit's length depends on the number of valid
// pixels of its block, which the speculative walk assumed to be 64: checked here against the mask.
template <class T>
__global__ void __launch_bounds__(FD_DWARPS * 32) k_dec_offsets(FastDecArgs a, const uint8_t* __restrict__ bits, uint32_t* __restrict__ blockOff) {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  __shared__ uint16_t sPosAll[FD_DWARPS * (FD_BATCH + 8)];
  __shared__ uint16_t sPairsAll[FD_DWARPS * FD_PTAB];
  __shared__ FdEntry sTab[FD_REG * FD_CAND];
  __shared__ uint32_t sTrue[(FD_REG + 1) * 3];
  __shared__ int sWhy;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int reg = blockIdx.x;
  const int sub0 = reg * a.subPerReg, nLocal = max(0, min(a.nSub, sub0 + a.subPerReg) - sub0);
  const int nBlocks = a.nTx * a.nTy;
  const int version = a.version;
  if (tid == 0) sWhy = 0;
  for (int i = tid; i < nLocal * FD_CAND; i += blockDim.x) sTab[i] = a.subTab[(size_t)sub0 * FD_CAND + i];
  __syncthreads();
  // ---- true entry (position, block index, candidate slot) of every sub-chunk of the region
  if (tid < 32) fdTrueEntries(a, reg, nLocal, nBlocks, sTab, sTrue, &sWhy, lane);
  __syncthreads();

  const int tailRaw = 1 + (a.nRows - (a.nTy - 1) * 8) * (a.nCols - (a.nTx - 1) * 8) * (int)sizeof(T);
  (void)tailRaw; (void)version;
  uint16_t* sPos = sPosAll + warp * (FD_BATCH + 8);
  uint16_t* sPairs = sPairsAll + warp * FD_PTAB;
  for (int ls = warp; ls < nLocal; ls += FD_DWARPS) {
    const uint32_t pos0 = sTrue[3 * ls], blk0 = sTrue[3 * ls + 1], slot = sTrue[3 * ls + 2];
    if (pos0 >= FD_DEAD - 1) continue;
    const int s = sub0 + ls;
    const uint32_t expectExit = sTrue[3 * ls + 3], expectBlk = sTrue[3 * ls + 4];
    const FdEntry me = sTab[ls * FD_CAND + slot];
    const int cnt = (int)me.count;
    bool fallback = false; unsigned why = 0;
    const int nPairs = fdLoadPairs<MAXU>(a.lens + ((size_t)s * FD_CAND + slot) * FD_LENS, pos0 - (uint32_t)s * FD_SUB, lane, sPairs);
    for (int batchBase = 0; batchBase < cnt; batchBase += FD_BATCH) {
      fdFillBatch(sPairs, nPairs, batchBase, cnt, me.exit - (uint32_t)s * FD_SUB, lane, sPos);
      const int nB = min(FD_BATCH, cnt - batchBase);
      for (int i = lane; i < nB; i += 32) {
        const uint32_t b = blk0 + (uint32_t)(batchBase + i);
        if (b >= (uint32_t)nBlocks) break;
        const uint32_t p = sPos[i], len = (uint32_t)sPos[i + 1] - p;
        const unsigned long long gp = (unsigned long long)s * FD_SUB + p;
        const unsigned flag = a.stream[gp];
        if ((flag & 3) == 0) {                                           // raw: 1 + (valid pixels of the block) * sizeof(T) bytes
          const int ty = (int)b / a.nTx, tx = (int)b - ty * a.nTx;
          const int h = min(8, a.nRows - ty * 8), w = min(8, a.nCols - tx * 8);
          int nv = 0;
          for (int rr = 0; rr < h; rr++) {
            const long long k0 = (long long)(ty * 8 + rr) * a.nCols + tx * 8;
            for (int c = 0; c < w; c++) nv += maskBit(bits, k0 + c) ? 1 : 0;
          }
          if (len != 1u + (uint32_t)nv * (uint32_t)sizeof(T)) { fallback = true; why |= 1024; }
        }
        blockOff[b] = (uint32_t)gp;
      }
      __syncwarp();
    }
    const uint32_t blkEnd = blk0 + (uint32_t)cnt;
    const bool haveNext = expectExit < FD_DEAD - 1;
    if (haveNext ? (blkEnd != expectBlk || me.exit != expectExit) : (blkEnd < (uint32_t)nBlocks)) { fallback = true; why |= 2048; }
    fallback = __any_sync(FULL, fallback);
    why = __reduce_or_sync(FULL, why);
    if (lane == 0 && fallback) atomicOr(a.status, DECF_FALLBACK | (int)why);
    __syncwarp();
  }
  if (tid == 0 && sWhy) atomicOr(a.status, DECF_FALLBACK | sWhy);
}

// ================= kernel 4: true entries of the region's sub-chunks, decode the blocks ================
template <class T>
__global__ void __launch_bounds__(FD_DWARPS * 32) k_dec_blocks(FastDecArgs a) {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  constexpr int BUFB = ((FD_SUB + MAXU + 64 + 15) / 16) * 16;       // per-warp staging of one sub-chunk (+ look-ahead)
  extern __shared__ __align__(16) uint8_t smemD[];                   // [FD_DWARPS][BUFB] stream staging
  __shared__ uint16_t sPosAll[FD_DWARPS * (FD_BATCH + 8)];
  __shared__ uint16_t sPairsAll[FD_DWARPS * FD_PTAB];
  __shared__ FdEntry sTab[FD_REG * FD_CAND];
  __shared__ uint32_t sTrue[(FD_REG + 1) * 3];
  __shared__ int sWhy;
  uint8_t* bufAll = smemD;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int reg = blockIdx.x;
  const int sub0 = reg * a.subPerReg, nLocal = max(0, min(a.nSub, sub0 + a.subPerReg) - sub0);
  const int nBlocks = a.nTx * a.nTy;
  const int version = a.version;
  if (tid == 0) sWhy = 0;
  for (int i = tid; i < nLocal * FD_CAND; i += blockDim.x) sTab[i] = a.subTab[(size_t)sub0 * FD_CAND + i];
  __syncthreads();
  // ---- true entry (position, block index, candidate slot) of every sub-chunk of the region
  if (tid < 32) fdTrueEntries(a, reg, nLocal, nBlocks, sTab, sTrue, &sWhy, lane);
  __syncthreads();

  // ---- decode: one warp per sub-chunk, 4 blocks at a time, 8 lanes per block (lane r = block row r)
  T* data = (T*)a.data;
  const bool vecOk = ((a.nCols * (int)sizeof(T)) % 16 == 0) && (((uintptr_t)data & 15) == 0);
  const int g = lane >> 3, r = lane & 7;
  uint8_t* buf = bufAll + (size_t)warp * BUFB;
  uint16_t* sPos = sPosAll + warp * (FD_BATCH + 8);
  uint16_t* sPairs = sPairsAll + warp * FD_PTAB;
  for (int ls = warp; ls < nLocal; ls += FD_DWARPS) {
    const uint32_t pos0 = sTrue[3 * ls], blk0 = sTrue[3 * ls + 1], slot = sTrue[3 * ls + 2];
    if (pos0 >= FD_DEAD - 1) continue;                               // dead chain (reported through sWhy) or past the end
    const int s = sub0 + ls;
    const uint32_t expectExit = sTrue[3 * ls + 3], expectBlk = sTrue[3 * ls + 4];
    const FdEntry me = sTab[ls * FD_CAND + slot];
    const int cnt = (int)me.count;
    // stage the sub-chunk: buf[d + i] = stream[s * FD_SUB + i]
    const unsigned long long start = (unsigned long long)s * FD_SUB;
    const uint8_t* gsub = a.stream + start;
    const int d = (int)((uintptr_t)gsub & 15);
    {
      const uint8_t* g0 = gsub - d;
      const long long avail = (long long)(a.streamLen - start) + d;
      for (int i = lane; i < BUFB / 16; i += 32) {
        uint4 x = make_uint4(0, 0, 0, 0);
        if ((long long)i * 16 < avail) x = __ldg((const uint4*)g0 + i);
        ((uint4*)buf)[i] = x;
      }
    }
    // block positions from the recorded unit lengths, FD_BATCH blocks at a time: sPos[i] = start of block batchBase + i
    const int nPairs = fdLoadPairs<MAXU>(a.lens + ((size_t)s * FD_CAND + slot) * FD_LENS, pos0 - (uint32_t)s * FD_SUB, lane, sPairs);
    __syncwarp();
    const uint8_t* sb = buf + d;
    const uint32_t* words = (const uint32_t*)buf;
    bool fallback = false; unsigned why = 0;
    for (int batchBase = 0; batchBase < cnt; batchBase += FD_BATCH) {
    fdFillBatch(sPairs, nPairs, batchBase, cnt, me.exit - (uint32_t)s * FD_SUB, lane, sPos);
    const int nB = min(FD_BATCH, cnt - batchBase);
    for (int i0 = 0; i0 < nB; i0 += 4) {
      const int i = i0 + g;
      const uint32_t b = blk0 + (uint32_t)(batchBase + i);
      if (i < nB && b < (uint32_t)nBlocks) {
        const int p = sPos[i], pNext = sPos[i + 1];
        const int ty = (int)b / a.nTx, tx = (int)b - ty * a.nTx;
        const int bi0 = ty * 8, bj0 = tx * 8;
        const int h = min(8, a.nRows - bi0), w = min(8, a.nCols - bj0), cells = h * w;
        T out[8];
        unsigned whyB = 0;
        const int len = fdDecodeBlockRow<T>(words, sb, d, p, version, tx & (version >= 5 ? 14 : 15), cells, h, w, r, a.invScale, a.zMax, out, whyB);
        if (whyB) { fallback = true; why |= whyB; }
        // the block, parsed with its true size, must end where the recorded chain continues
        if (!fallback && p + len != pNext) { fallback = true; why |= 1024; }
        if (!fallback && r < h) fdStoreRow<T>(data + (size_t)(bi0 + r) * a.nCols + bj0, out, w, vecOk);
      }
    }
    __syncwarp();
    }
    // the serial parse continues in the next sub-chunk where the resolution assumed, with the block index it assumed;
    // or the stream's blocks end in here
    const uint32_t blkEnd = blk0 + (uint32_t)cnt;
    const bool haveNext = expectExit < FD_DEAD - 1;
    if (haveNext ? (blkEnd != expectBlk || me.exit != expectExit) : (blkEnd < (uint32_t)nBlocks)) { fallback = true; why |= 2048; }
    fallback = __any_sync(FULL, fallback);
    why = __reduce_or_sync(FULL, why);
    if (lane == 0 && fallback) atomicOr(a.status, DECF_FALLBACK | (int)why);
    __syncwarp();
  }
  if (tid == 0 && sWhy) atomicOr(a.status, DECF_FALLBACK | sWhy);
}

template <class T> inline size_t fastDecodeWalkSmem() {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  return (size_t)FD_REG * FD_SUB + ((MAXU + 64 + 15) / 16) * 16 + 32;
}
template <class T> inline size_t fastDecodeBlocksSmem() {
  constexpr int MAXU = 1 + 64 * (int)sizeof(T);
  constexpr int BUFB = ((FD_SUB + MAXU + 64 + 15) / 16) * 16;
  return (size_t)FD_DWARPS * BUFB;
}

}  // namespace lerc
