// lerc_decode_stream.cuh -- single-kernel decoder of a Lerc2 micro-block stream for the headline raster shape:
// every pixel valid, nDepth == 1, 8x8 micro-blocks (included by lerc_decode.cu after lerc_decode_fast.cuh, whose unit
// parsers and block-row decoder it uses).
//
// The stream has no index: a block's length is only known from its own header bytes (ReadTile, Lerc2.cpp:2025-2230;
// BitStuffer2::Decode, BitStuffer2.cpp:159-258), so the reference walks it serially.  Here one CTA takes one 16 KB chunk
// of the stream (by ticket, in stream order) and does everything for it from ONE staged copy:
//
//   stage      the chunk (+ look-ahead for the unit that straddles its end) is copied into shared memory by the copy
//              engine: one cp.async.bulk (TMA, SASS UBLKCP) from the 16-byte aligned address below the chunk, completion
//              on an mbarrier; the last partial 16 bytes of the stream are read byte-wise (nothing is read past the blob)
//   guess      the chunk is cut into 16 sub-chunks of 1 KB; every warp scans the head windows of four of them (the true
//              chain must enter a sub-chunk inside its first MAXU bytes), 32 byte positions per step, for the first position
//              that reads as the headers of two full bit-stuffed blocks in a row (mode 1, one-byte count of 64, consecutive
//              integrity bits); windows without one (flat regions) take the first position from which DS_HOPS units parse with
//              consistent integrity bits (Lerc2.cpp:2045).  A wrong guess is a position inside the previous unit whose chain
//              merges into the true chain or dies - it differs from the true chain in its first few hops only.
//   walk       warp 0, one LANE per sub-chunk: hop from header to header to the end of the sub-chunk, recording the positions
//   publish    the chunk's speculative exit (where its last chain leaves the chunk) for the next chunk; the previous
//              chunk's exit is this chunk's TRUE entry (chunk 0 starts at 0)
//   patch      one thread walks from the true entry until it hits a recorded position (usually at once), sub-chunk by
//              sub-chunk: from there on the recorded chain IS the serial parse.  The chunk's exact block count follows.
//              If the exit published before turns out wrong, or anything does not parse, the kernel raises DSF_FALLBACK.
//   look-back  decoupled look-back over the chunks' block counts (lerc_lookback.cuh): index of the chunk's first block
//   decode     8 lanes per block, one lane per block row: the block is re-parsed with its true size, must carry the
//              integrity bits of its column and must end exactly where the next unit starts; unpack, z = offset + q * 2
//              maxZError in fp64 without contraction, min(z, zMax), cast, 128-bit stores (Lerc2.cpp:2145-2160)
//   checksum   Fletcher-32 partial sums of the chunk's bytes from the same shared-memory copy; the last CTA to finish
//              compares the blob's checksum (Lerc2.cpp:1037-1064)
//
// By induction over the chunks (chunk 0's entry is certain; every chunk checks that the exit it published is the exit of
// the parse from its true entry) the result is the serial parse, or DSF_FALLBACK is raised and the caller runs the
// multi-kernel speculative decoder / the general decoder, which decide what is malformed exactly like the reference.
#pragma once
#include "lerc_tma.cuh"
#include "lerc_lookback.cuh"
#include "lerc_fletcher.cuh"

namespace lerc {

constexpr int DS_CHUNK = 16384;                 // stream bytes per CTA
constexpr int DS_SUBS = 16;                     // sub-chunks per chunk, one lane of warp 0 each
constexpr int DS_SUB = DS_CHUNK / DS_SUBS;
constexpr int DS_LIST_OFFS = 1032;              // ... in the offsets-only instantiation (masked rasters: runs of one-byte units of empty blocks)
constexpr int DS_LIST = 264;                    // recorded positions per sub-chunk; more units than that in 1 KB (flat regions: 1..3-byte blocks) -> DSF_FALLBACK
constexpr int DS_PATCH = 48;                    // hops the patch walk may need before it joins the recorded chain
constexpr int DS_HOPS = 4;                      // units a head-window position must parse to become the guess (windows without a strict candidate)
constexpr int DS_THREADS = 128;
enum : unsigned int { DSF_FALLBACK = 8, DSF_CHECKSUM = 2, DSF_REPORTED = 0x80000000u };

struct StreamDecResult {                        // device, zero-initialised per call
  unsigned int ticket, done, status, pad;
  unsigned long long fletA, fletD;
};

struct StreamDecArgs {
  const uint8_t* stream; unsigned long long streamLen;
  int nRows, nCols, nTx, nTy, version;
  uint32_t nTxMagic;                            // floor(2^32 / nTx) + 1 (nTx >= 2)
  double invScale, zMax;                        // 2 * maxZError ; header zMax
  void* data;
  int nChunks;
  int chunkBegin;                               // this launch: chunks chunkBegin .. chunkBegin + gridDim.x - 1 (a stream that arrives in strips is decoded strip by strip)
unsigned long long* hostEnd;                  // mapped host word or nullptr
  unsigned int* hostStatus;                     // mapped host word or nullptr: the final status word | DSF_REPORTED, written by the last CTA of the last launch
    uint32_t* blockOff;                           // OFFS instantiation: [nBlocks] stream offset of every block
  unsigned int* ticket;                         // this launch's ticket counter (zero at launch)
  unsigned long long* exitState;                // [nChunks] 0 = not yet, else (stream offset where the chunk's chain leaves it) + 1; bit 63: no chain
  unsigned long long* cntState;                 // [nChunks] look-back words over the chunks' block counts
  unsigned long long* groupState;               // [ceil(nChunks / 32)]
  unsigned long long* groupAcc;                 // [ceil(nChunks / 32)]
  StreamDecResult* res;
  long long regionOff, regionLen;               // checksum region: offset of stream[0] in it, its length (blobSize - 14)
  unsigned long long prefA, prefD;              // Fletcher partial sums of the region's bytes before the stream (host)
  uint32_t expectChecksum; int haveChecksum;
};

template <class T> struct DecStream {
  static constexpr int MAXU = 1 + 64 * (int)sizeof(T);                  // longest unit: the raw 8x8 block
  static constexpr int LA = ((MAXU + 64 + 15) / 16) * 16;               // look-ahead behind the chunk
  static constexpr int BUFB = 16 + DS_CHUNK + LA;                       // alignment slack | chunk | look-ahead
  static constexpr int SMEM = BUFB + DS_SUBS * DS_LIST * 2;
  static constexpr int SMEM_OFFS = BUFB + DS_SUBS * DS_LIST_OFFS * 2;
};

// b / nTx and b % nTx for b < 2^32 with the precomputed magic (exact: the estimate is never more than one too high)
__device__ __forceinline__ void dsDivMod(uint32_t b, int nTx, uint32_t magic, int& q, int& r) {
  uint32_t qq = nTx == 1 ? b : __umulhi(b, magic);
  if (nTx != 1 && qq * (uint32_t)nTx > b) qq--;
  q = (int)qq; r = (int)(b - qq * (uint32_t)nTx);
}

// Bytes of a block offset per offset-type code (Lerc2.h:528-542, offsetTypeFromCode + dtSize as a nibble table; 0 = no such type)
template <class T> struct DsOszTable {
  static constexpr int DT = PixelTraits<T>::code;
  static constexpr uint32_t v = DT <= DT_Byte ? 0x1111u : DT == DT_Short ? 0x0112u : DT == DT_UShort ? 0x0012u : DT == DT_Int ? 0x1224u
                              : DT == DT_UInt ? 0x0124u : DT == DT_Float ? 0x1124u : 0x2448u;
};
// Header of a FULL bit-stuffed block (mode 1, no LUT, one-byte count == 64) at byte pointer q (shared memory)?  Straight-line code, three
// byte loads: cheap enough to test every byte position of a head window.  len = the unit's length, pat = its integrity bits.
template <class T>
__device__ __forceinline__ bool dsStrict(const uint8_t* __restrict__ q, int version, int& len, int& pat) {
  const uint32_t flag = q[0];
  const int osz = (int)((DsOszTable<T>::v >> (4 * (flag >> 6))) & 15u);
  const uint32_t b = q[1 + osz], n = q[2 + osz];
  pat = fdPattern(flag, version);
  len = 3 + osz + 8 * (int)(b & 31);
  return (flag & 3) == 1 && !(version >= 5 && (flag & 4)) && osz != 0 && (b & 0xe0) == 0x80 && (b & 31) != 0 && n == 64;
}

// OFFS: the same boundary discovery for rasters with a mask (nDepth == 1): no pixels, no checksum; the stream offset of every block goes to
// a.blockOff and the general block decoder (k_tiles_decode) does the rest, after k_verify_offsets has checked every unit with
// its true valid count against the chain.  Empty blocks are one-byte units, hence the longer position lists.
template <class T, bool OFFS = false>
__global__ void __launch_bounds__(DS_THREADS, OFFS ? 4 : 6) k_decode_stream(StreamDecArgs a) {
  using C = DecStream<T>;
  constexpr int LIST = OFFS ? DS_LIST_OFFS : DS_LIST;
  constexpr int PATCH = OFFS ? C::MAXU + 7 : DS_PATCH;           // (offsets: a run of one-byte units may fill the whole head window in front of the guess)
  constexpr int MAXU = C::MAXU;
  extern __shared__ __align__(16) uint8_t dsSmem[];
  uint8_t* buf = dsSmem;                                   // buf[d + i] = stream[start + i]
  uint16_t* sListAll = (uint16_t*)(dsSmem + C::BUFB);      // [DS_SUBS][LIST] recorded positions (relative to the chunk start)
  __shared__ uint16_t sPatch[DS_SUBS][PATCH];
  __shared__ int sGuess[DS_SUBS], sCnt[DS_SUBS], sExit[DS_SUBS], sDead[DS_SUBS];
  __shared__ int sFirst[DS_SUBS], sNPatch[DS_SUBS], sPre[DS_SUBS + 1], sTrueExit[DS_SUBS];
  __shared__ __align__(8) uint64_t sBar;
  __shared__ int sChunk, sOk, sTotal;
  __shared__ unsigned long long sBlk0;
  __shared__ unsigned long long sFA[DS_THREADS / 32], sFD[DS_THREADS / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int version = a.version;
  const int nBlocks = a.nTx * a.nTy;
  const int tailRaw = 1 + (a.nRows - (a.nTy - 1) * 8) * (a.nCols - (a.nTx - 1) * 8) * (int)sizeof(T);

  // ---- ticket, copy engine
  if (tid == 0) {
    const int c = a.chunkBegin + (int)atomicAdd(a.ticket, 1u);
    sChunk = c;
    const unsigned long long start = (unsigned long long)c * DS_CHUNK;
    const long long left = (long long)(a.streamLen - start);
    const int avail = (int)min((long long)(DS_CHUNK + C::LA), left);
    const uint8_t* g = a.stream + start;
    const int d = (int)((uintptr_t)g & 15);
    const int total = d + avail, bulk = total & ~15;
    mbarInit(&sBar, 1);
    if (bulk) { mbarExpectTx(&sBar, (uint32_t)bulk); bulkLoad(buf, g - d, (uint32_t)bulk, &sBar); mbarSimCopiesDone(&sBar); }
    for (int i = bulk; i < total; i++) buf[i] = (g - d)[i];          // the stream's last partial 16 bytes
  }
  __syncthreads();
  const int c = sChunk;
  const unsigned long long start = (unsigned long long)c * DS_CHUNK;
  const long long left = (long long)(a.streamLen - start);
  const int avail = (int)min((long long)(DS_CHUNK + C::LA), left);
  const int chunkLen = (int)min((long long)DS_CHUNK, left);
  const int d = (int)((uintptr_t)(a.stream + start) & 15);
  for (int i = d + avail + tid; i < C::BUFB; i += DS_THREADS) buf[i] = 0;
  if ((d + avail) & ~15) mbarWait(&sBar, 0);
  __syncthreads();
  const uint32_t* words = (const uint32_t*)buf;
  const uint8_t* sb = buf + d;
  const int nSubs = (chunkLen + DS_SUB - 1) / DS_SUB;
  const int testable = (int)min((long long)(DS_CHUNK + C::LA - 32), left);   // positions below it have a full window staged

  // one unit at chunk-relative position p (speculative: an 8x8 block is assumed); 0 = does not parse
  auto hopPlain = [&](int p, int& pat) -> int {
    int len;
    if (dsStrict<T>(sb + p, version, len, pat) && (long long)len <= left - p) return len;      // the common unit, three byte loads
    return fdHopLen<T>(fdWindow(words, (uint32_t)(d + p)), sb + p, version, left - p, tailRaw, pat);
  };
  // `hops` more units parse from q with consecutive integrity bits.  level 0: none of them may look like a raw unit; level 1: one raw
  // unit on the way is resolved by trying its sizes with a level-0 look-ahead behind it; level 2: a raw unit is taken on its
  // integrity bits alone.  (The bytes of raw pixels often look like short units; an encoder's empty-block byte has zero top bits.)
  auto looksRaw = [&](uint32_t f) { return (f & 3) == 0 && !(version >= 5 && (f & 4)); };
  auto chainFrom = [&](int q, int patPrev, int hops, int level) -> bool {
    for (int hI = 0; hI < hops; hI++) {
      if ((long long)q >= left || q >= testable) return true;
      const uint32_t f = sb[q];
      if (looksRaw(f)) {
        const int pf = fdPattern(f, version);
        if (level == 0 || !fdFollows(patPrev, pf, version)) return false;
        if (level == 2) return true;
        for (int n2 = 1; n2 <= 64; n2++) {                              // level 1: sizes of the nested raw unit
          const long long q2 = (long long)q + 1 + (long long)n2 * (int)sizeof(T);
          if (q2 > left) return false;
          if (q2 == left || q2 >= testable) return true;
          // (inlined level-0 look-ahead, 6 units)
          int qq = (int)q2, pp = pf; bool ok0 = true;
          for (int h2 = 0; h2 < 6 && ok0; h2++) {
            if ((long long)qq >= left || qq >= testable) break;
            const uint32_t f2 = sb[qq];
            if (looksRaw(f2) || ((f2 & 3) == 2 && (f2 & 0xc0))) { ok0 = false; break; }
            int pt; const int l2 = hopPlain(qq, pt);
            if (l2 <= 0 || !fdFollows(pp, pt, version)) ok0 = false; else { qq += l2; pp = pt; }
          }
          if (ok0) return true;
        }
        return false;
      }
      if ((f & 3) == 2 && (f & 0xc0)) return false;
      int pt; const int len = hopPlain(q, pt);
      if (len <= 0 || !fdFollows(patPrev, pt, version)) return false;
      q += len; patPrev = pt;
    }
    return true;
  };
  auto hopLen = [&](int p, int& pat) -> int {
    if constexpr (OFFS) {
      // A raw unit of a partly valid block is 1 + n * sizeof(T) bytes for the block's n valid pixels, which only the block's index would
      // tell.  The smallest n behind which eight more units parse with consecutive integrity bits is taken (raw wins for small n only),
      // strictest look-ahead first; k_verify_offsets checks the result against the mask.
      const uint32_t flag = sb[p];
      if (looksRaw(flag)) {
        pat = fdPattern(flag, version);
        for (int level = 0; level < 3; level++)
          for (int n = 1; n <= 64; n++) {
            const int len = 1 + n * (int)sizeof(T);
            if ((long long)len > left - p) break;
            if ((long long)(p + len) == left) return len;              // ends the stream
            if (p + len >= testable) break;
            if (chainFrom(p + len, pat, 8, level)) return len;
          }
#ifdef LERC_CUSIM
        if (std::getenv("DS_DEBUG3")) { std::fprintf(stderr, "      raw trial failed: chunk %d p %d bytes", c, p); for (int k = 0; k < 24; k++) std::fprintf(stderr, " %02x", sb[p + k]); std::fprintf(stderr, "\n"); }
#endif
        return 0;
      }
    }
    return hopPlain(p, pat);
  };

  // ---- guess: every warp scans the head windows of four sub-chunks
  for (int s = warp; s < nSubs; s += DS_THREADS / 32) {
    const int subStart = s * DS_SUB;
    int guess = -1;
    if (c == 0 && s == 0) guess = 0;                                  // the stream starts with block 0
    else {
      const int headEnd = min(subStart + MAXU, testable);
      for (int base = subStart; base < headEnd && guess < 0; base += 32) {
        const int p = base + lane;
        // two full bit-stuffed blocks in a row with consecutive integrity bits (one alone can be faked by the bytes of a block's offset)
        int len = 0, pat1 = 0;
        bool hit = p < headEnd && dsStrict<T>(sb + p, version, len, pat1) && (long long)len <= left - p;
        if (__any_sync(FULL, hit)) {
          if (hit && (long long)len != left - p) {                       // (a block that ends the stream needs no successor)
            int len2 = 0, pat2 = 0;
            hit = p + len < testable && dsStrict<T>(sb + p + len, version, len2, pat2) && (long long)len2 <= left - p - len && fdFollows(pat1, pat2, version);
          }
        }
        const unsigned m = __ballot_sync(FULL, hit);
        if (m) guess = base + __ffs(m) - 1;
      }
      if (guess < 0) {
        // no full bit-stuffed block starts in the window (flat region, edge blocks, raw blocks): general units, DS_HOPS of them with
        // consistent integrity bits; raw units carry no redundancy, so a position whose first unit is not raw is preferred
        int reserve = -1;
        for (int base = subStart; base < headEnd && guess < 0; base += 32) {
          const int p = base + lane;
          bool alive = p < headEnd, firstRaw = false;
          int q = p, pat = 0;
          for (int hop = 0; hop < DS_HOPS; hop++) {
            if (alive && q < testable) {
              int np;
              const int len = hopPlain(q, np);                            // (guessing never resolves raw units by trial: too many bytes look like one)
              if (len <= 0 || (hop > 0 && !fdFollows(pat, np, version))) alive = false;
              else { if (hop == 0) firstRaw = len == MAXU; q += len; pat = np; }
            }
          }
          const unsigned m1 = __ballot_sync(FULL, alive && !firstRaw), m2 = __ballot_sync(FULL, alive);
          if (m1) guess = base + __ffs(m1) - 1;
          else if (m2 && reserve < 0) reserve = base + __ffs(m2) - 1;
        }
        if (guess < 0) guess = reserve;
      }
    }
    if (lane == 0) sGuess[s] = guess;
  }
  __syncthreads();

  if (warp == 0) {
    // ---- walk: lane s hops through sub-chunk s from its guess, recording the positions
    {
      const int s = lane, subEnd = min((s + 1) * DS_SUB, chunkLen);
      uint16_t* list = sListAll + s * LIST;
      int n = 0, dead = 0;
      int guess = s < nSubs ? sGuess[s] : -1, p = guess;
      const int headEnd = min(s * DS_SUB + MAXU, testable);
      for (int attempt = 0; attempt < 64 && guess >= 0; attempt++) {
        n = 0; dead = 0; p = guess;
        int pat = 0;
        while (p < subEnd) {
          if (n >= LIST) { dead = 1; break; }
          int np;
          const int len = hopLen(p, np);
          if (len <= 0 || (n > 0 && !fdFollows(pat, np, version))) { dead = 1; break; }
          list[n++] = (uint16_t)p;
          p += len; pat = np;
        }
        // a chain that dies was a wrong guess (the true chain of a well-formed stream does not die): the next position of the head
        // window from which a unit parses is tried (rare; the other lanes wait)
        if (!dead || (c == 0 && s == 0)) break;
        int q = guess + 1;
        for (; q < headEnd; q++) { int np; if (hopPlain(q, np) > 0) break; }
        if (q >= headEnd) break;
        guess = q;
      }
      if (s < nSubs) { sCnt[s] = n; sExit[s] = p; sDead[s] = dead; }
      __syncwarp();

      // ---- publish the speculative exit, take the true entry (lane 0), then every lane joins its sub-chunk's true entry - the exit of
      // the sub-chunk before it - with its recorded chain: usually its first position; else a few hops from the entry until a recorded
      // position is hit.  A lane whose chain was a stranger (never joined) changes its exit; its successors are redone.
      volatile unsigned long long* ex = a.exitState;
      const int sL = nSubs - 1;
      const int guessL = __shfl_sync(FULL, guess, sL), deadL = __shfl_sync(FULL, dead, sL), exitL = __shfl_sync(FULL, p, sL);
      const bool spec = guessL >= 0 && !deadL;
      const unsigned long long specExit = spec ? start + (unsigned long long)exitL : 0;
      long long entry0 = 0;
      bool ok = true;
      if (lane == 0) {
        if (spec) ex[c] = specExit + 1;
        if (c > 0) {
          unsigned long long v;
          while ((v = ex[c - 1]) == 0) __nanosleep(100);
          if (v >> 63) ok = false; else entry0 = (long long)(v - 1) - (long long)start;
          if (entry0 < 0) ok = false;
        }
      }
      ok = __shfl_sync(FULL, ok ? 1 : 0, 0) != 0;
      const int exitSpec = p;
      int curExit = exitSpec, first = n, npatch = 0, cs = 0, prevEntry = -1;
      bool bad = false, pending = true;
      for (int iter = 0; iter < 12 && pending; iter++) {
        int e = __shfl_up_sync(FULL, curExit, 1);
        if (lane == 0) e = (int)entry0;
        const bool need = s < nSubs && e != prevEntry;
#ifdef LERC_CUSIM
        if (std::getenv("DS_DEBUG2") && need) std::fprintf(stderr, "      iter %d lane %d e %d prev %d n %d list0 %d\n", iter, lane, e, prevEntry, n, n ? (int)list[0] : -1);
#endif
        if (need) {
          prevEntry = e;
          int q = e, cursor = 0;
          bool joined = false;
          npatch = 0; bad = false;
          while (q < subEnd) {
            while (cursor < n && (int)list[cursor] < q) cursor++;
            if (cursor < n && (int)list[cursor] == q) { joined = true; break; }
            if (npatch == PATCH || q >= testable) { bad = true; break; }
            int pat;
            const int len = hopLen(q, pat);
#ifdef LERC_CUSIM
            if (len <= 0 && std::getenv("DS_DEBUG3")) { std::fprintf(stderr, "      patch hop failed: chunk %d sub %d q %d npatch %d bytes", c, s, q, npatch); for (int k = 0; k < 24; k++) std::fprintf(stderr, " %02x", sb[q + k]); std::fprintf(stderr, "\n"); }
#endif
            if (len <= 0) { bad = true; break; }
            sPatch[s][npatch++] = (uint16_t)q;
            q += len;
          }
          if (joined) { if (dead) bad = true; first = cursor; cs = npatch + n - cursor; curExit = exitSpec; }   // (dead: the true chain runs into the unit that did not parse)
          else if (bad) curExit = exitSpec;                           // (an entry that does not parse came from a stranger before this lane: wait for its correction)
          else { first = n; cs = npatch; curExit = q; }
        }
        pending = __any_sync(FULL, need);
      }
      if (pending || __any_sync(FULL, bad && s < nSubs)) ok = false;
      // block counts -> exclusive prefix over the sub-chunks
      int inc = s < nSubs ? cs : 0;
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) { const int o = __shfl_up_sync(FULL, inc, m); if (lane >= m) inc += o; }
      int total = __shfl_sync(FULL, inc, 31);
      const unsigned long long finalExit = start + (unsigned long long)__shfl_sync(FULL, curExit, sL);
      if (ok && spec && finalExit != specExit) ok = false;            // the next chunk may have started from a wrong entry
      if (s < DS_SUBS) { sFirst[s] = first; sNPatch[s] = npatch; sPre[s] = inc - (s < nSubs ? cs : 0); sTrueExit[s] = curExit; }
      if (lane == 31) sPre[DS_SUBS] = total;
#ifdef LERC_CUSIM
      if (std::getenv("DS_DEBUG")) {
        if (lane == 0) std::fprintf(stderr, "[ds] chunk %d ok %d entry %lld spec %d specExit %llu final %llu start %llu chunkLen %d left %lld total %d\n", c, (int)ok, entry0, (int)spec, specExit, finalExit, start, chunkLen, left, total);
        if (s < nSubs) std::fprintf(stderr, "   sub %d guess %d cnt %d exit %d dead %d | first %d npatch %d cs %d trueExit %d bad %d\n", s, guess, n, exitSpec, dead, first, npatch, cs, curExit, (int)bad);
      }
#endif
      if (lane == 0) {
        if (ok && !spec) ex[c] = finalExit + 1;
        if (!ok) {
          atomicOr(&a.res->status, DSF_FALLBACK);
          if (!spec) ex[c] = 1ull << 63;
        }
      }
      if (!ok) total = 0;
      if (lane == 0) { sOk = ok ? 1 : 0; sTotal = total; lookbackPublish(a.cntState, a.groupAcc, c, (unsigned long long)total); }
    }
    __syncwarp();
    // ---- index of the chunk's first block
    const int totalW = sTotal;
    const unsigned long long blk0 = lookbackExclusive(a.cntState, a.groupAcc, a.groupState, c, (unsigned long long)totalW, lane);
    if (lane == 0) {
      sBlk0 = blk0;
      if (c == a.chunkBegin + (int)gridDim.x - 1 && a.hostEnd) *(volatile unsigned long long*)a.hostEnd = blk0 + (unsigned long long)totalW;   // (mapped host memory: blocks behind this launch)
      if (c == a.nChunks - 1 && sOk && blk0 + (unsigned long long)totalW != (unsigned long long)nBlocks) atomicOr(&a.res->status, DSF_FALLBACK | 2048);
    }
  }

  // ---- Fletcher-32 partial sums of the chunk's bytes (warps 1..7 while warp 0 walks and looks back)
  unsigned long long fa = 0, fd = 0;
  if (a.haveChecksum && warp > 0) {
    const unsigned par = (unsigned)((a.regionOff + (long long)(uintptr_t)a.stream) & 1);
    const int nGroups = (d + chunkLen + 15) >> 4;
    for (int j = tid - 32; j < nGroups; j += DS_THREADS - 32) {
      const uint4 x = ((const uint4*)buf)[j];
      uint32_t o[4] = {x.x, x.y, x.z, x.w};
      const int r = j * 16 - d;                                         // chunk-relative byte of the group's first byte
      if (r < 0 || r + 16 > chunkLen) {
#pragma unroll
        for (int q = 0; q < 16; q++) if (r + q < 0 || r + q >= chunkLen) o[q >> 2] &= ~(0xffu << (8 * (q & 3)));
      }
      const long long r0 = a.regionOff + (long long)start + r;          // region offset of the group's first byte; r0 & 1 == par
      const uint32_t w0 = (uint32_t)((unsigned long long)(r0 - par) >> 1) % 65535u;
      uint32_t S, S1;
      if (par) fletcherChunk<1>(o, S, S1); else fletcherChunk<0>(o, S, S1);
      fa += S; fd += (unsigned long long)w0 * S + S1;
    }
  }
  __syncthreads();
  const bool ok = sOk != 0;

  // ---- decode: 8 lanes per block (lane r = block row r); lane group j takes the blocks of sub-chunk j one after the other
  if (ok && OFFS) {
    // ---- offsets only: every recorded unit is one block (nDepth == 1)
    const unsigned long long blk0 = sBlk0;
    bool fallback = false;
    for (int s = tid >> 3; s < nSubs; s += DS_THREADS / 8) {
      const int npS = sNPatch[s], firstS = sFirst[s], cs = sPre[s + 1] - sPre[s];
      const uint16_t* list = sListAll + s * LIST;
      const unsigned long long b0 = blk0 + (unsigned long long)sPre[s];
      if (cs > 0 && b0 + (unsigned long long)cs > (unsigned long long)nBlocks) { fallback = true; continue; }
      for (int i = tid & 7; i < cs; i += 8) {
        const int pos = i < npS ? (int)sPatch[s][i] : (int)list[firstS + i - npS];
        a.blockOff[b0 + (unsigned long long)i] = (uint32_t)(start + (unsigned long long)pos);
      }
    }
    if (fallback) atomicOr(&a.res->status, DSF_FALLBACK | 4096);
  }
  if (ok && !OFFS) {
    T* data = (T*)a.data;
    const bool vecOk = ((a.nCols * (int)sizeof(T)) % 16 == 0) && (((uintptr_t)data & 15) == 0);
    const unsigned long long blk0 = sBlk0;
    const int r = tid & 7;
    const float zMaxF = (float)a.zMax;
    bool fallback = false; unsigned why = 0;
    for (int s = tid >> 3; s < nSubs && !fallback; s += DS_THREADS / 8) {
    const int npS = sNPatch[s], firstS = sFirst[s], cs = sPre[s + 1] - sPre[s];
    const uint16_t* list = sListAll + s * LIST;
    int p = cs > 0 ? (npS > 0 ? (int)sPatch[s][0] : (int)list[firstS]) : 0;
    int ty = 0, tx = 0;
    if (cs > 0) {
      const unsigned long long b0 = blk0 + (unsigned long long)sPre[s];
      if (b0 + (unsigned long long)cs > (unsigned long long)nBlocks) { fallback = true; why |= 4096; }
      else dsDivMod((uint32_t)b0, a.nTx, a.nTxMagic, ty, tx);
    }
    for (int i = 0; i < cs && !fallback; i++) {
      const int i1 = i + 1;
      const int pNext = i1 < cs ? (i1 < npS ? (int)sPatch[s][i1] : (int)list[firstS + i1 - npS]) : sTrueExit[s];
      const int bi0 = ty * 8, bj0 = tx * 8;
      const int h = min(8, a.nRows - bi0), w = min(8, a.nCols - bj0);
      const int patExpect = tx & (version >= 5 ? 14 : 15);
      bool done = false;
      if constexpr (std::is_same<T, float>::value) {
        // hot path: full 8x8 float block, bit-stuffed with at most 16 bits, one-byte count
        const FdWin win = fdWindow(words, (uint32_t)(d + p));
        const uint32_t flag = (uint32_t)win.lo & 0xff;
        const int tc = (int)(flag >> 6), osz = tc == 0 ? 4 : (tc == 1 ? 2 : 1);
        const uint32_t bq = fdByte(win, 1 + osz), n = fdByte(win, 2 + osz);
        const int nb = (int)(bq & 31);
        if ((flag & 3) == 1 && !(version >= 5 && (flag & 4)) && fdPattern(flag, version) == patExpect && h == 8 && w == 8 &&
            (bq & 0xe0) == 0x80 && n == 64 && nb >= 1 && nb <= 16 && p + 3 + osz + 8 * nb == pNext) {
          const uint32_t ob = (uint32_t)(win.lo >> 8);
          const double offset = tc == 0 ? (double)__uint_as_float(ob) : (tc == 1 ? (double)(int16_t)(uint16_t)ob : (double)(uint8_t)ob);
          const FdWin rw = fdWindow(words, (uint32_t)(d + p + 3 + osz + r * nb));
          // value k sits at bit k * nb of the row's 128 bits: values 0..3 in the low 64 bits (4 nb <= 64), values 4..7 in the 64 bits from bit 4 nb
          const uint32_t m1 = (1u << nb) - 1;
          const int s4 = 4 * nb;
          const unsigned long long g1 = s4 == 64 ? rw.hi : ((rw.lo >> s4) | (rw.hi << (64 - s4)));
          uint32_t qv[8];
          qv[0] = (uint32_t)rw.lo & m1; qv[1] = (uint32_t)(rw.lo >> nb) & m1; qv[2] = (uint32_t)(rw.lo >> (2 * nb)) & m1; qv[3] = (uint32_t)(rw.lo >> (3 * nb)) & m1;
          qv[4] = (uint32_t)g1 & m1; qv[5] = (uint32_t)(g1 >> nb) & m1; qv[6] = (uint32_t)(g1 >> (2 * nb)) & m1; qv[7] = (uint32_t)(g1 >> (3 * nb)) & m1;
          float out[8];
          // (T)min(z, zMax) == min((T)z, (T)zMax): the conversion is monotonic (Lerc2.cpp:2160; zMax is finite, checked by the host)
#pragma unroll
          for (int kk = 0; kk < 8; kk++) out[kk] = fminf((float)__dadd_rn(offset, __dmul_rn((double)qv[kk], a.invScale)), zMaxF);
          fdStoreRow<float>((float*)data + (size_t)(bi0 + r) * a.nCols + bj0, out, 8, vecOk);
          done = true;
        }
      }
      if (!done) {
        T out[8];
        unsigned whyB = 0;
        const int len = fdDecodeBlockRow<T>(words, sb, d, p, version, patExpect, h * w, h, w, r, a.invScale, a.zMax, out, whyB);
        if (whyB) { fallback = true; why |= whyB; }
        else if (p + len != pNext) { fallback = true; why |= 1024; }    // parsed with its true size the block must end where the chain continues
        else if (r < h) fdStoreRow<T>(data + (size_t)(bi0 + r) * a.nCols + bj0, out, w, vecOk);
      }
      p = pNext;
      if (++tx == a.nTx) { tx = 0; ty++; }
    }
    }
    if (fallback) atomicOr(&a.res->status, DSF_FALLBACK | why);
  }

  // ---- checksum partials of the CTA; the last CTA compares
  fa %= 65535ull; fd %= 65535ull;
#pragma unroll
  for (int m = 16; m; m >>= 1) { fa += __shfl_xor_sync(FULL, fa, m); fd += __shfl_xor_sync(FULL, fd, m); }
  if (lane == 0) { sFA[warp] = fa; sFD[warp] = fd; }
  __syncthreads();
  if (tid == 0) {
    unsigned long long A = 0, D = 0;
    for (int i = 0; i < DS_THREADS / 32; i++) { A += sFA[i]; D += sFD[i]; }
    if (A | D) { atomicAdd(&a.res->fletA, A); atomicAdd(&a.res->fletD, D % 65535ull); }
    __threadfence();
    const unsigned int prev = atomicAdd(&a.res->done, 1u);
    if (prev == (unsigned int)a.nChunks - 1) {
      __threadfence();
      unsigned int bits = 0;
      if (a.haveChecksum) {
        const unsigned long long tA = (*(volatile unsigned long long*)&a.res->fletA + a.prefA) % 65535ull;
        const unsigned long long tD = (*(volatile unsigned long long*)&a.res->fletD % 65535ull + a.prefD) % 65535ull;
        if (fletcherFinishFast(tA, tD, a.regionLen) != a.expectChecksum) bits = DSF_CHECKSUM;
      }
      const unsigned int all = atomicOr(&a.res->status, bits) | bits;
      if (a.hostStatus) *(volatile unsigned int*)a.hostStatus = all | DSF_REPORTED;          // (mapped host memory: no copy behind the row copies)
    }
  }
}

}  // namespace lerc
