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
//   guess      one warp per 2 KB sub-chunk: the lowest byte position of the sub-chunk's head window (the true chain must
//              enter inside the first MAXU bytes) from which DS_HOPS consecutive units parse with consistent integrity
//              bits (Lerc2.cpp:2045).  A wrong guess is a position inside the previous unit whose chain has merged into
//              the true chain - it differs from the true chain in its first few hops only.
//   walk       lane 0 of every warp hops from header to header to the end of its sub-chunk, recording the positions
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
constexpr int DS_SUBS = 8;                      // sub-chunks per chunk, one warp each
constexpr int DS_SUB = DS_CHUNK / DS_SUBS;
constexpr int DS_LIST = DS_SUB + 8;             // recorded positions per sub-chunk (1-byte units fill it with DS_SUB)
constexpr int DS_PATCH = 64;                    // hops the patch walk may need before it joins the recorded chain
constexpr int DS_HOPS = 4;                      // units a head-window position must parse to become the guess
constexpr int DS_THREADS = DS_SUBS * 32;
enum { DSF_FALLBACK = 8, DSF_CHECKSUM = 2 };

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
};

// b / nTx and b % nTx for b < 2^32 with the precomputed magic (exact: the estimate is never more than one too high)
__device__ __forceinline__ void dsDivMod(uint32_t b, int nTx, uint32_t magic, int& q, int& r) {
  uint32_t qq = nTx == 1 ? b : __umulhi(b, magic);
  if (nTx != 1 && qq * (uint32_t)nTx > b) qq--;
  q = (int)qq; r = (int)(b - qq * (uint32_t)nTx);
}

template <class T>
__global__ void __launch_bounds__(DS_THREADS, 4) k_decode_stream(StreamDecArgs a) {
  using C = DecStream<T>;
  constexpr int MAXU = C::MAXU;
  extern __shared__ __align__(16) uint8_t dsSmem[];
  uint8_t* buf = dsSmem;                                   // buf[d + i] = stream[start + i]
  uint16_t* sListAll = (uint16_t*)(dsSmem + C::BUFB);      // [DS_SUBS][DS_LIST] recorded positions (relative to the chunk start)
  __shared__ uint16_t sPatch[DS_SUBS][DS_PATCH];
  __shared__ int sGuess[DS_SUBS], sCnt[DS_SUBS], sExit[DS_SUBS], sDead[DS_SUBS];
  __shared__ int sFirst[DS_SUBS], sNPatch[DS_SUBS], sPre[DS_SUBS + 1], sTrueExit[DS_SUBS];
  __shared__ __align__(8) uint64_t sBar;
  __shared__ int sChunk, sOk, sTotal;
  __shared__ unsigned long long sBlk0;
  __shared__ unsigned long long sFA[DS_SUBS], sFD[DS_SUBS];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int version = a.version;
  const int nBlocks = a.nTx * a.nTy;
  const int tailRaw = 1 + (a.nRows - (a.nTy - 1) * 8) * (a.nCols - (a.nTx - 1) * 8) * (int)sizeof(T);

  // ---- ticket, copy engine
  if (tid == 0) {
    const int c = (int)atomicAdd(&a.res->ticket, 1u);
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
  auto hopLen = [&](int p, int& pat) -> int {
    return fdHopLen<T>(fdWindow(words, (uint32_t)(d + p)), sb + p, version, left - p, tailRaw, pat);
  };

  // ---- guess (whole warp) + walk (lane 0) of sub-chunk `warp`
  if (warp < nSubs) {
    const int s = warp, subStart = s * DS_SUB, subEnd = min(subStart + DS_SUB, chunkLen);
    uint16_t* list = sListAll + s * DS_LIST;
    // A guess whose chain dies before the end of the sub-chunk was a wrong one (the true chain of a well-formed stream never
    // dies): the search resumes behind it, a few times.
    int searchFrom = subStart;
    for (int attempt = 0; attempt < 6; attempt++) {
      int guess = -1;
      if (c == 0 && s == 0) guess = 0;                                // the stream starts with block 0
      else {
        // Raw units carry no redundancy (any byte with zero mode bits "is" a raw block of MAXU bytes), so a position whose first
        // unit is not raw is preferred wherever it lies in the head window; the lowest other survivor is the reserve.
        const int headEnd = min(subStart + MAXU, testable);
        int reserve = -1;
        for (int base = searchFrom; base < headEnd && guess < 0; base += 32) {
          const int p = base + lane;
          bool alive = p < headEnd, firstRaw = false;
          int q = p, pat = 0;
          for (int hop = 0; hop < DS_HOPS; hop++) {
            if (alive && q < testable) {
              int np;
              const int len = hopLen(q, np);
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
      int dead = 0;
      if (lane == 0) {
        int n = 0, p = guess, pat = 0;
        if (guess >= 0) {
          while (p < subEnd) {
            if (n >= DS_LIST) { dead = 1; break; }
            int np;
            const int len = hopLen(p, np);
            if (len <= 0 || (n > 0 && !fdFollows(pat, np, version))) { dead = 1; break; }
            list[n++] = (uint16_t)p;
            p += len; pat = np;
          }
        }
        sGuess[s] = guess; sCnt[s] = n; sExit[s] = p; sDead[s] = dead;
      }
      dead = __shfl_sync(FULL, dead, 0);
      if (!dead || guess < 0 || (c == 0 && s == 0)) break;
      searchFrom = guess + 1;
    }
  }
  __syncthreads();

  // ---- publish the speculative exit, take the true entry, patch, count (thread 0)
  if (tid == 0) {
    volatile unsigned long long* ex = a.exitState;
    const int sL = nSubs - 1;
    const bool spec = nSubs > 0 && sGuess[sL] >= 0 && !sDead[sL];
    const unsigned long long specExit = spec ? start + (unsigned long long)sExit[sL] : 0;
    if (spec) ex[c] = specExit + 1;
    bool ok = nSubs > 0;
    long long p = 0;
    if (c > 0) {
      unsigned long long v;
      while ((v = ex[c - 1]) == 0) __nanosleep(100);
      if (v >> 63) ok = false; else p = (long long)(v - 1) - (long long)start;
      if (p < 0) ok = false;
    }
    int total = 0;
    for (int s = 0; s < nSubs && ok; s++) {
      const int subEnd = min((s + 1) * DS_SUB, chunkLen);
      const uint16_t* list = sListAll + s * DS_LIST;
      const int cnt = sCnt[s];
      int cursor = 0, np = 0;
      bool joined = false;
      while (p < subEnd) {
        while (cursor < cnt && (long long)list[cursor] < p) cursor++;
        if (cursor < cnt && (long long)list[cursor] == p) { joined = true; break; }
        if (np == DS_PATCH || p >= testable) { ok = false; break; }
        int pat;
        const int len = hopLen((int)p, pat);
        if (len <= 0) { ok = false; break; }
        sPatch[s][np++] = (uint16_t)p;
        p += len;
      }
      if (!ok) break;
      int cs = np, first = cnt;
      if (joined) {
        if (sDead[s]) { ok = false; break; }                          // the true chain runs into the unit that did not parse
        first = cursor; cs += cnt - cursor; p = sExit[s];
      }
      sFirst[s] = first; sNPatch[s] = np; sPre[s] = total; sTrueExit[s] = (int)p;
      total += cs;
    }
    if (ok) {
      for (int s = nSubs; s <= DS_SUBS; s++) sPre[s] = total;
      const unsigned long long finalExit = start + (unsigned long long)p;
      if (spec) { if (finalExit != specExit) ok = false; }            // the next chunk may have started from a wrong entry
      else ex[c] = finalExit + 1;
    }
#ifdef LERC_CUSIM
    if (std::getenv("DS_DEBUG")) {
      std::fprintf(stderr, "[ds] chunk %d ok %d p %lld spec %d specExit %llu start %llu chunkLen %d left %lld\n", c, (int)ok, p, (int)spec, specExit, start, chunkLen, left);
      for (int s = 0; s < nSubs; s++) std::fprintf(stderr, "   sub %d guess %d cnt %d exit %d dead %d | first %d npatch %d pre %d trueExit %d\n", s, sGuess[s], sCnt[s], sExit[s], sDead[s], sFirst[s], sNPatch[s], sPre[s], sTrueExit[s]);
    }
#endif
    if (!ok) {
      total = 0;
      atomicOr(&a.res->status, DSF_FALLBACK);
      if (!spec) ex[c] = 1ull << 63;
    }
    sOk = ok ? 1 : 0; sTotal = total;
    lookbackPublish(a.cntState, a.groupAcc, c, (unsigned long long)total);
  }
  __syncthreads();
  const int total = sTotal;
  const bool ok = sOk != 0;

  // ---- index of the chunk's first block
  if (warp == 0) {
    const unsigned long long blk0 = lookbackExclusive(a.cntState, a.groupAcc, a.groupState, c, (unsigned long long)total, lane);
    if (lane == 0) {
      sBlk0 = blk0;
      if (c == a.nChunks - 1 && ok && blk0 + (unsigned long long)total != (unsigned long long)nBlocks) atomicOr(&a.res->status, DSF_FALLBACK | 2048);
    }
  }

  // ---- Fletcher-32 partial sums of the chunk's bytes (warps 1..7 start while warp 0 looks back)
  unsigned long long fa = 0, fd = 0;
  if (a.haveChecksum) {
    const unsigned par = (unsigned)((a.regionOff + (long long)(uintptr_t)a.stream) & 1);
    const int nGroups = (d + chunkLen + 15) >> 4;
    for (int j = tid; j < nGroups; j += DS_THREADS) {
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

  // ---- decode: 8 lanes per block (lane r = block row r), 32 blocks per sweep
  if (ok) {
    T* data = (T*)a.data;
    const bool vecOk = ((a.nCols * (int)sizeof(T)) % 16 == 0) && (((uintptr_t)data & 15) == 0);
    const unsigned long long blk0 = sBlk0;
    const int r = tid & 7;
    bool fallback = false; unsigned why = 0;
    for (int k = tid >> 3; k < total; k += DS_THREADS / 8) {
      int s = 0;
#pragma unroll
      for (int j = 1; j < DS_SUBS; j++) s += (k >= sPre[j]) ? 1 : 0;
      const int i = k - sPre[s], npS = sNPatch[s], firstS = sFirst[s], cs = sPre[s + 1] - sPre[s];
      const uint16_t* list = sListAll + s * DS_LIST;
      const int p = i < npS ? (int)sPatch[s][i] : (int)list[firstS + i - npS];
      const int i1 = i + 1;
      const int pNext = i1 < cs ? (i1 < npS ? (int)sPatch[s][i1] : (int)list[firstS + i1 - npS]) : sTrueExit[s];
      const unsigned long long b = blk0 + (unsigned long long)k;
      if (b >= (unsigned long long)nBlocks) { fallback = true; why |= 4096; continue; }
      int ty, tx;
      dsDivMod((uint32_t)b, a.nTx, a.nTxMagic, ty, tx);
      const int bi0 = ty * 8, bj0 = tx * 8;
      const int h = min(8, a.nRows - bi0), w = min(8, a.nCols - bj0), cells = h * w;
      T out[8];
      unsigned whyB = 0;
      const int len = fdDecodeBlockRow<T>(words, sb, d, p, version, tx & (version >= 5 ? 14 : 15), cells, h, w, r, a.invScale, a.zMax, out, whyB);
      if (whyB) { fallback = true; why |= whyB; }
      else if (p + len != pNext) { fallback = true; why |= 1024; }      // parsed with its true size the block must end where the chain continues
      else if (r < h) fdStoreRow<T>(data + (size_t)(bi0 + r) * a.nCols + bj0, out, w, vecOk);
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
    for (int i = 0; i < DS_SUBS; i++) { A += sFA[i]; D += sFD[i]; }
    if (A | D) { atomicAdd(&a.res->fletA, A); atomicAdd(&a.res->fletD, D % 65535ull); }
    __threadfence();
    const unsigned int prev = atomicAdd(&a.res->done, 1u);
    if (prev == (unsigned int)a.nChunks - 1 && a.haveChecksum) {
      __threadfence();
      const unsigned long long tA = (*(volatile unsigned long long*)&a.res->fletA + a.prefA) % 65535ull;
      const unsigned long long tD = (*(volatile unsigned long long*)&a.res->fletD % 65535ull + a.prefD) % 65535ull;
      if (fletcherFinish(tA, tD, a.regionLen) != a.expectChecksum) atomicOr(&a.res->status, DSF_CHECKSUM);
    }
  }
}

}  // namespace lerc
