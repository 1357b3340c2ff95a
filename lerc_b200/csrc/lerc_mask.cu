// lerc_mask.cu -- validity-mask kernels (byte mask <-> MSB-first bit mask, NaN folding, byte RLE),
// Fletcher-32 and a few small utilities.  Reference citations relative to /root/reference/src/LercLib.
#include "lerc_device.cuh"
#include "lerc_kernels.h"
#include <cub/device/device_scan.cuh>
#include <algorithm>

namespace lerc {

// ------------------------------------------------------------------------------------------------
// byte mask (+ NaN test for float types) -> bit mask                 Lerc.cpp:959-975, :1420-1468
// One thread produces one mask byte (8 pixels).  Pad bits of the last byte stay 1 (SetAllValid then
// clear, Lerc.cpp:967-972).  counters[0] += number of valid pixels; counters[1] |= flags.
template <class T>
__global__ void k_mask_build(const T* __restrict__ data, const uint8_t* __restrict__ bytes, long long nPix, int nDepth,
                             uint8_t* __restrict__ bits, int* __restrict__ counters) {
  const long long nBytes = (nPix + 7) >> 3;
  int valid = 0, flags = 0;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nBytes; b += (long long)gridDim.x * blockDim.x) {
    unsigned out = 0;
    for (int j = 0; j < 8; j++) {
      const long long k = b * 8 + j;
      bool v = true;
      if (k < nPix) {
        if (bytes) v = bytes[k] != 0;
        if (PixelTraits<T>::isFloat && v) {
          int bad = 0;
          const T* px = data + k * nDepth;
          for (int m = 0; m < nDepth; m++) bad += isNaNVal(px[m]) ? 1 : 0;
          if (bad == nDepth) { v = false; flags |= MASKF_MODIFIED; }
          else if (bad > 0) flags |= MASKF_MIXED_NAN;
        }
        valid += v ? 1 : 0;
      }
      out |= (v ? 1u : 0u) << (7 - j);
    }
    bits[b] = (uint8_t)out;
  }
  valid = warpSum(valid);
  flags = __reduce_or_sync(FULL, flags);
  if ((threadIdx.x & 31) == 0) {
    if (valid) atomicAdd(&counters[0], valid);
    if (flags) atomicOr(&counters[1], flags);
  }
}

template <class T>
void launchMaskBuild(Context* ctx, const void* dData, const uint8_t* dBytes, long long nPix, int nDepth, uint8_t* dBits, int* dCounters) {
  const long long nBytes = (nPix + 7) >> 3;
  int grid = (int)std::min<long long>((nBytes + 255) / 256, 148 * 16);
  LERC_LAUNCH(ctx, k_mask_build<T>, grid, 256, 0, (const T*)dData, dBytes, nPix, nDepth, dBits, dCounters);
}
#define INST(T) template void launchMaskBuild<T>(Context*, const void*, const uint8_t*, long long, int, uint8_t*, int*);
INST(int8_t) INST(uint8_t) INST(int16_t) INST(uint16_t) INST(int32_t) INST(uint32_t) INST(float) INST(double)
#undef INST

// bit mask -> byte mask (Lerc.cpp:979-995)
__global__ void k_bits_to_bytes(const uint8_t* __restrict__ bits, long long nPix, uint8_t* __restrict__ bytes) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nPix; k += (long long)gridDim.x * blockDim.x)
    bytes[k] = maskBit(bits, k) ? 1 : 0;
}
void launchBitsToBytes(Context* ctx, const uint8_t* dBits, long long nPix, uint8_t* dBytes) {
  int grid = (int)std::min<long long>((nPix + 255) / 256, 148 * 32);
  LERC_LAUNCH(ctx, k_bits_to_bytes, grid, 256, 0, dBits, nPix, dBytes);
}

__global__ void k_fill_bytes(uint8_t* p, long long n, uint8_t v) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) p[k] = v;
}

// masks of consecutive bands equal?  (Lerc.cpp:999-1010; compared as validity, see DESIGN.md "Deviations")
__global__ void k_bits_differ(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long nPix, int* __restrict__ flag) {
  const long long nBytes = (nPix + 7) >> 3;
  int diff = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nBytes; i += (long long)gridDim.x * blockDim.x) {
    unsigned x = a[i] ^ b[i];
    if (i == nBytes - 1 && (nPix & 7)) x &= 0xffu << (8 - (nPix & 7));     // ignore pad bits
    diff |= x ? 1 : 0;
  }
  if (__any_sync(FULL, diff) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
void launchBitsDiffer(Context* ctx, const uint8_t* a, const uint8_t* b, long long nPix, int* dFlag) {
  const long long nBytes = (nPix + 7) >> 3;
  int grid = (int)std::min<long long>((nBytes + 255) / 256, 148 * 8);
  LERC_LAUNCH(ctx, k_bits_differ, grid, 256, 0, a, b, nPix, dFlag);
}

// number of valid pixels in each run of 1024 pixels (128 mask bytes); one warp per chunk
__global__ void k_chunk_valid_counts(const uint8_t* __restrict__ bits, long long nPix, int nChunks, uint32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  for (int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nChunks; c += gridDim.x * (blockDim.x >> 5)) {
    int n = 0;
    for (int j = 0; j < 4; j++) {
      const long long byteIdx = (long long)c * 128 + lane * 4 + j, k0 = byteIdx * 8;
      if (k0 < nPix) {
        unsigned b = bits[byteIdx];
        if (k0 + 8 > nPix) b &= 0xffu << (8 - (nPix - k0));
        n += __popc(b);
      }
    }
    n = warpSum(n);
    if (lane == 0) counts[c] = (uint32_t)n;
  }
}
void launchChunkValidCounts(Context* ctx, const uint8_t* dBits, long long nPix, int nChunks, uint32_t* dCounts) {
  int grid = std::min((nChunks + 7) / 8, 148 * 8);
  LERC_LAUNCH(ctx, k_chunk_valid_counts, grid, 256, 0, dBits, nPix, nChunks, dCounts);
}

// ------------------------------------------------------------------------------------------------
// byte RLE of the bit mask                                                          RLE.cpp:32-331
// Token stream: int16 count; count > 0: that many literal bytes follow; count < 0: one byte follows,
// repeated -count times; -32768 ends the stream.  A maximal run of equal bytes is coded as a repeat
// iff it is >= 5 long and starts more than 5 bytes before the end of the array (RLE.cpp:74-79); all
// counts are capped at 32767 (RLE.cpp:98-107).  One warp walks the runs; lanes measure run lengths
// with ballots and copy literal stretches cooperatively.  dst == nullptr only sizes.
__global__ void __launch_bounds__(32) k_rle_encode(const uint8_t* __restrict__ src, long long n, uint8_t* __restrict__ dst, uint32_t* __restrict__ sizeOut) {
  // The run scan is a serial dependency chain; it reads the mask through a 16 KB shared-memory window (refilled with
  // independent 128-bit loads) so that a step costs a shared-memory access, not a DRAM round trip.
  constexpr int WIN = 16384;
  __shared__ __align__(16) uint8_t win[WIN + 16];
  const int lane = threadIdx.x;
  long long wBase = 0; bool haveWin = false;               // window holds src[wBase, wBase + WIN); src + wBase is 16-byte aligned
  const int mis = (int)((uintptr_t)src & 15);
  auto fill = [&](long long i) {                          // make src[i] the first 16-byte chunk of the window
    const long long a0 = ((i + mis) & ~15ll) - mis;       // src index (may be negative by < 16) whose address is 16-byte aligned
    for (int c = lane; c < WIN / 16; c += 32) {
      const long long s0 = a0 + (long long)c * 16;
      uint4 x = make_uint4(0, 0, 0, 0);
      if (s0 < n && s0 + 16 > 0) x = *(const uint4*)(src + s0);          // whole chunk lies inside the allocation granule of src
      ((uint4*)win)[c] = x;
    }
    wBase = a0; haveWin = true;
    __syncwarp();
  };
  auto rd = [&](long long i) -> unsigned { return win[i - wBase]; };
  long long pos = 0, lit = 0, out = 0;
  auto putCount = [&](int c) {
    if (dst && lane == 0) { dst[out] = (uint8_t)(c & 0xff); dst[out + 1] = (uint8_t)((c >> 8) & 0xff); }
    out += 2;
  };
  auto flushLiterals = [&](long long a, long long b) {
    while (a < b) {
      const long long c = (b - a > 32767) ? 32767 : (b - a);
      putCount((int)c);
      if (dst) for (long long i = lane; i < c; i += 32) dst[out + i] = src[a + i];
      out += c; a += c;
    }
  };
  while (pos < n) {
    if (!haveWin || pos < wBase || pos + 64 > wBase + WIN) fill(pos);
    const unsigned v = rd(pos);
    long long run = 1;
    for (;;) {                                   // extend the run 32 bytes at a time
      const long long i = pos + run + lane;
      if (pos + run + 32 > wBase + WIN) { __syncwarp(); fill(pos + run); }
      const unsigned m = __ballot_sync(FULL, i < n && rd(i) == v);
      const int ones = (m == FULL) ? 32 : (__ffs(~m) - 1);
      run += ones;
      if (ones < 32) break;
    }
    if (run >= 5 && pos + 5 < n) {
      flushLiterals(lit, pos);
      long long left = run;
      while (left) {
        const long long c = left > 32767 ? 32767 : left;
        putCount(-(int)c);
        if (dst && lane == 0) dst[out] = (uint8_t)v;
        out += 1; left -= c;
      }
      lit = pos + run;
    }
    pos += run;
  }
  flushLiterals(lit, n);
  putCount(-32768);
  if (lane == 0) *sizeOut = (uint32_t)out;
}
// ---- the same token stream in parallel (masks of more than a few KB): run starts, literal-span starts and the ends of both are
// found with prefix scans instead of a walk.
//   k_rle_runs    every byte that starts a maximal run of equal bytes gets a marker {position, "repeat" bit}; the repeat rule
//                 (run >= 5 and more than 5 bytes before the end, RLE.cpp:74-79) needs only the four bytes behind the run start
//   (max-scan)    -> every byte knows its run's start and repeat bit
//   k_rle_spans   start markers of repeat runs / literal spans (maximal stretches of non-repeat bytes) in stream order and end
//                 markers in REVERSED order, packed into one 64-bit element
//   (max-scan of both halves at once) -> every byte knows where its repeat run / literal span starts and ends
//   k_rle_sizes   output bytes contributed by every byte: a piece of at most 32767 (RLE.cpp:98-107) starts every 32767 bytes of a
//                 repeat run (3 bytes) or literal span (2 bytes + the literals themselves)
//   (sum-scan)    -> output offsets
//   k_rle_write   headers with the piece's count (what is left of the run / span, capped), bytes, terminator
struct U32Max { __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };
struct U32x2 { uint32_t a, b; };
struct U32x2Max { __host__ __device__ U32x2 operator()(const U32x2& x, const U32x2& y) const { U32x2 r; r.a = x.a > y.a ? x.a : y.a; r.b = x.b > y.b ? x.b : y.b; return r; } };

__global__ void k_rle_runs(const uint8_t* __restrict__ src, uint32_t n, uint32_t* __restrict__ mark) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint8_t v = src[i];
    uint32_t m = 0;
    if (i == 0 || src[i - 1] != v) {
      const bool rep = i + 5 < n && src[i + 1] == v && src[i + 2] == v && src[i + 3] == v && src[i + 4] == v;
      m = ((i + 1) << 1) | (rep ? 1u : 0u);
    }
    mark[i] = m;
  }
}
__global__ void k_rle_spans(const uint8_t* __restrict__ src, uint32_t n, const uint32_t* __restrict__ run, U32x2* __restrict__ se) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t r = run[i], p = (r >> 1) - 1;
    const bool rep = r & 1u;
    const bool runEnd = i == n - 1 || src[i + 1] != src[i];
    bool start, end;
    if (rep) { start = i == p; end = runEnd; }
    else {
      start = i == 0 || (i == p && (run[i - 1] & 1u));                    // the byte before belongs to a repeat run
      end = i == n - 1 || (runEnd && (run[i + 1] & 1u));                  // the next run is a repeat run
    }
    se[i].a = start ? i + 1 : 0;
    se[n - 1 - i].b = end ? (n - 1 - i) + 1 : 0;
  }
}
__global__ void k_rle_sizes(uint32_t n, const uint32_t* __restrict__ run, const U32x2* __restrict__ se, uint32_t* __restrict__ contrib) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const bool rep = run[i] & 1u;
    const uint32_t k = i - (se[i].a - 1);
    const bool piece = k % 32767u == 0;
    contrib[i] = rep ? (piece ? 3u : 0u) : (piece ? 3u : 1u);
  }
}
__global__ void k_rle_write(const uint8_t* __restrict__ src, uint32_t n, const uint32_t* __restrict__ run, const U32x2* __restrict__ se,
                            const uint32_t* __restrict__ off, uint8_t* __restrict__ dst, uint32_t* __restrict__ sizeOut) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const bool rep = run[i] & 1u;
    const uint32_t st = se[i].a - 1, en = n - 1 - (se[n - 1 - i].b - 1);
    const uint32_t k = i - st;
    const bool piece = k % 32767u == 0;
    const uint32_t o = off[i];
    if (dst) {
      if (piece) {
        const uint32_t left = en - i + 1, c = left > 32767u ? 32767u : left;
        const uint32_t cnt = rep ? (0x10000u - c) & 0xffffu : c;            // int16 count, little endian; negative = repeat
        dst[o] = (uint8_t)cnt; dst[o + 1] = (uint8_t)(cnt >> 8); dst[o + 2] = src[i];
      } else if (!rep) dst[o] = src[i];
    }
    if (i == n - 1) {
      const uint32_t total = o + (rep ? (piece ? 3u : 0u) : (piece ? 3u : 1u));
      if (dst) { dst[total] = 0x00; dst[total + 1] = 0x80; }               // -32768 ends the stream (RLE.cpp:250)
      *sizeOut = total + 2;
    }
  }
}

void launchRleEncode(Context* ctx, const uint8_t* dSrc, long long n, uint8_t* dDst, uint32_t* dSize) {
  if (n < 4096 || n >= (1ll << 30)) { LERC_LAUNCH(ctx, k_rle_encode, 1, 32, 0, dSrc, n, dDst, dSize); return; }
  const uint32_t nn = (uint32_t)n;
  uint32_t* dRun = (uint32_t*)ctx->arena.alloc(4 * (size_t)nn);
  uint32_t* dOff = (uint32_t*)ctx->arena.alloc(4 * (size_t)nn);
  U32x2* dSe = (U32x2*)ctx->arena.alloc(8 * (size_t)nn);
  size_t t1 = 0, t2 = 0, t3 = 0;
  cub::DeviceScan::InclusiveScan(nullptr, t1, dRun, dRun, U32Max(), (int)nn, ctx->stream);
  cub::DeviceScan::InclusiveScan(nullptr, t2, dSe, dSe, U32x2Max(), (int)nn, ctx->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, t3, dOff, dOff, (int)nn, ctx->stream);
  const size_t tb = std::max(std::max(t1, t2), std::max(t3, (size_t)16));
  void* tmp = ctx->arena.alloc(tb);
  if (!dRun || !dOff || !dSe || !tmp) { LERC_LAUNCH(ctx, k_rle_encode, 1, 32, 0, dSrc, n, dDst, dSize); return; }
  const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 16);
  LERC_LAUNCH(ctx, k_rle_runs, grid, 256, 0, dSrc, nn, dRun);
  { LaunchScope scope(ctx, "cub::InclusiveScan<max>"); size_t b = tb; cub::DeviceScan::InclusiveScan(tmp, b, dRun, dRun, U32Max(), (int)nn, ctx->stream); ctx->kernelLaunches += 2; }
  LERC_LAUNCH(ctx, k_rle_spans, grid, 256, 0, dSrc, nn, dRun, dSe);
  { LaunchScope scope(ctx, "cub::InclusiveScan<max2>"); size_t b = tb; cub::DeviceScan::InclusiveScan(tmp, b, dSe, dSe, U32x2Max(), (int)nn, ctx->stream); ctx->kernelLaunches += 2; }
  LERC_LAUNCH(ctx, k_rle_sizes, grid, 256, 0, nn, dRun, dSe, dOff);
  { LaunchScope scope(ctx, "cub::ExclusiveSum<u32>"); size_t b = tb; cub::DeviceScan::ExclusiveSum(tmp, b, dOff, dOff, (int)nn, ctx->stream); ctx->kernelLaunches += 2; }
  LERC_LAUNCH(ctx, k_rle_write, grid, 256, 0, dSrc, nn, dRun, dSe, dOff, dDst, dSize);
}

// RLE.cpp:298-331.  status[0] = 1 on success, 0 on malformed input.
__global__ void __launch_bounds__(32) k_rle_decode(const uint8_t* __restrict__ src, long long srcLen, uint8_t* __restrict__ dst, long long dstLen, int* __restrict__ status) {
  // token headers are a serial dependency chain: read them through a shared-memory window (see k_rle_encode)
  constexpr int WIN = 8192;
  __shared__ __align__(16) uint8_t win[WIN + 16];
  const int lane = threadIdx.x;
  long long wBase = 0; bool haveWin = false;
  const int mis = (int)((uintptr_t)src & 15);
  auto fill = [&](long long i) {
    const long long a0 = ((i + mis) & ~15ll) - mis;
    for (int c = lane; c < WIN / 16; c += 32) {
      const long long s0 = a0 + (long long)c * 16;
      uint4 x = make_uint4(0, 0, 0, 0);
      if (s0 < srcLen && s0 + 16 > 0) x = *(const uint4*)(src + s0);
      ((uint4*)win)[c] = x;
    }
    wBase = a0; haveWin = true;
    __syncwarp();
  };
  long long ip = 0, op = 0;
  int ok = 1;
  for (;;) {
    if (ip + 2 > srcLen) { ok = 0; break; }
    if (!haveWin || ip < wBase || ip + 4 > wBase + WIN) { __syncwarp(); fill(ip); }
    const int c = (int)(int16_t)(win[ip - wBase] | (win[ip + 1 - wBase] << 8));
    ip += 2;
    if (c == -32768) break;
    const long long cnt = c <= 0 ? -c : c, take = c > 0 ? cnt : 1;
    if (ip + take + 2 > srcLen || op + cnt > dstLen) { ok = 0; break; }
    if (c > 0) for (long long i = lane; i < cnt; i += 32) dst[op + i] = src[ip + i];
    else { const uint8_t v = win[ip - wBase]; for (long long i = lane; i < cnt; i += 32) dst[op + i] = v; }
    ip += take; op += cnt;
  }
  if (lane == 0) *status = ok;
}
// ---- the same in two steps for masks of more than a few KB: one warp hops over the token HEADERS only (the chain is serial, but a
// hop is a few shared-memory reads, whatever the token's length) and lists the tokens; then every thread fills 16 output bytes, finding
// its token by binary search.  Verdict and bytes written are those of k_rle_decode (RLE.cpp:298-331).
struct RleTok { uint32_t ip, op; int32_t c; };        // payload position, output position, count (> 0 literal bytes, <= 0 one byte repeated)

__global__ void __launch_bounds__(32) k_rle_tokens(const uint8_t* __restrict__ src, long long srcLen, long long dstLen, RleTok* __restrict__ tok, uint32_t maxTok,
                                                   uint32_t* __restrict__ nTokOut, int* __restrict__ status) {
  // The header chain is serial.  Per 8 KB window all lanes first work out, for EVERY byte position, the token that would start there
  // (its length in the stream and its count), so that the hop itself is two shared-memory reads and an add.
  constexpr int WIN = 8192;
  __shared__ __align__(16) uint8_t win[WIN + 16];
  __shared__ uint16_t sLen[WIN];                   // 0: terminator, or a header that does not lie wholly inside the window
  __shared__ int16_t sCnt[WIN];
  constexpr int LISTCAP = WIN / 3 + 1;
  __shared__ uint16_t sList[LISTCAP];              // positions of the window's tokens
  const int lane = threadIdx.x;
  long long wBase = 0; bool haveWin = false;
  const int mis = (int)((uintptr_t)src & 15);
  auto fill = [&](long long i) {
    const long long a0 = ((i + mis) & ~15ll) - mis;
    for (int c = lane; c < WIN / 16; c += 32) {
      const long long s0 = a0 + (long long)c * 16;
      uint4 x = make_uint4(0, 0, 0, 0);
      if (s0 < srcLen && s0 + 16 > 0) x = *(const uint4*)(src + s0);
      ((uint4*)win)[c] = x;
    }
    wBase = a0; haveWin = true;
    __syncwarp();
    for (int p = lane; p < WIN; p += 32) {
      int len = 0, c = 0;
      if (p + 2 <= WIN) {
        c = (int)(int16_t)(win[p] | (win[p + 1] << 8));
        if (c != -32768) len = 2 + (c > 0 ? c : 1);
      }
      sLen[p] = (uint16_t)len; sCnt[p] = (int16_t)c;
    }
    __syncwarp();
  };
  long long ip = 0, op = 0;
  uint32_t n = 0;
  int ok = 1;
  for (;;) {
    if (ip + 2 > srcLen) { ok = 0; break; }
    if (!haveWin || ip < wBase || ip + 4 > wBase + WIN) { __syncwarp(); fill(ip); }
    // ---- the window's tokens: positions by a bare hop loop (two shared-memory reads and an add per token), then counts, output
    // positions (warp scan), bounds and records for 32 tokens at a time
    {
      int rel = (int)(ip - wBase), k = 0;
      while (rel + 4 <= WIN && k < LISTCAP) {
        const int len = sLen[rel];
        if (len == 0) break;
        if (lane == 0) sList[k] = (uint16_t)rel;
        k++; rel += len;
      }
      __syncwarp();
      bool stop = false;
      for (int base = 0; base < k && !stop; base += 32) {
        const int i = base + lane;
        const bool have = i < k;
        const int r = have ? (int)sList[i] : 0;
        const int c = have ? (int)sCnt[r] : 0, len = have ? (int)sLen[r] : 0;
        const long long cnt = c <= 0 ? -c : c;
        long long inc = cnt;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) { const long long o = __shfl_up_sync(FULL, inc, m); if (lane >= m) inc += o; }
        const long long myOp = op + inc - cnt, myIp = wBase + r;
        const bool bad = have && (myIp + len + 2 > srcLen || myOp + cnt > dstLen || (unsigned long long)n + (unsigned)(i - base) >= maxTok);
        const unsigned mb = __ballot_sync(FULL, bad), mh = __ballot_sync(FULL, have);
        const int nGood = mb ? __ffs(mb) - 1 : __popc(mh);                 // tokens of this batch before the first one that needs the exact path
        if (lane < nGood) { RleTok t; t.ip = (uint32_t)(myIp + 2); t.op = (uint32_t)myOp; t.c = c; tok[n + lane] = t; }
        if (nGood > 0) {
          const long long endIp = __shfl_sync(FULL, myIp + len, nGood - 1), endOp = __shfl_sync(FULL, myOp + cnt, nGood - 1);
          ip = endIp; op = endOp; n += (uint32_t)nGood;
        }
        if (mb) stop = true;
      }
      if (ip + 2 > srcLen) { ok = 0; break; }
      if (ip < wBase || ip + 4 > wBase + WIN) { __syncwarp(); fill(ip); }
    }
    // ---- one token the exact way (terminator, bounds, list full)
    const int c = (int)(int16_t)(win[ip - wBase] | (win[ip + 1 - wBase] << 8));
    ip += 2;
    if (c == -32768) break;
    const long long cnt = c <= 0 ? -c : c, take = c > 0 ? cnt : 1;
    if (ip + take + 2 > srcLen || op + cnt > dstLen || n >= maxTok) { ok = 0; break; }
    if (lane == 0) { RleTok t; t.ip = (uint32_t)ip; t.op = (uint32_t)op; t.c = c; tok[n] = t; }
    n++;
    ip += take; op += cnt;
  }
  if (lane == 0) { RleTok t; t.ip = 0; t.op = (uint32_t)op; t.c = 0; tok[n] = t; *nTokOut = n; *status = ok; }   // (sentinel: where the output ends)
}

__global__ void k_rle_expand(const uint8_t* __restrict__ src, const RleTok* __restrict__ tok, const uint32_t* __restrict__ nTokIn, uint8_t* __restrict__ dst) {
  const uint32_t nTok = *nTokIn;
  if (nTok == 0) return;
  const uint32_t total = tok[nTok].op;
  for (uint32_t o0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16u; o0 < total; o0 += gridDim.x * blockDim.x * 16u) {
    uint32_t lo = 0, hi = nTok - 1;                                   // last token with op <= o0
    while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (tok[mid].op <= o0) lo = mid; else hi = mid - 1; }
    uint32_t t = lo;
    RleTok cur = tok[t];
    uint32_t end = tok[t + 1].op;
    const uint32_t oEnd = min(o0 + 16u, total);
    for (uint32_t o = o0; o < oEnd; o++) {
      while (o >= end) { t++; cur = tok[t]; end = tok[t + 1].op; }
      dst[o] = cur.c > 0 ? src[cur.ip + (o - cur.op)] : src[cur.ip];
    }
  }
}

void launchRleDecode(Context* ctx, const uint8_t* dSrc, long long srcLen, uint8_t* dDst, long long dstLen, int* dStatus) {
  if (srcLen < 2048 || srcLen >= (1ll << 31) || dstLen >= (1ll << 31)) { LERC_LAUNCH(ctx, k_rle_decode, 1, 32, 0, dSrc, srcLen, dDst, dstLen, dStatus); return; }
  const uint32_t maxTok = (uint32_t)(srcLen / 3 + 2);
  RleTok* dTok = (RleTok*)ctx->arena.alloc(sizeof(RleTok) * ((size_t)maxTok + 2));
  uint32_t* dN = (uint32_t*)ctx->arena.alloc(16);
  if (!dTok || !dN) { LERC_LAUNCH(ctx, k_rle_decode, 1, 32, 0, dSrc, srcLen, dDst, dstLen, dStatus); return; }
  LERC_LAUNCH(ctx, k_rle_tokens, 1, 32, 0, dSrc, srcLen, dstLen, dTok, maxTok, dN, dStatus);
  const int grid = (int)std::min<long long>((dstLen / 16 + 255) / 256 + 1, 148 * 16);
  LERC_LAUNCH(ctx, k_rle_expand, grid, 256, 0, dSrc, dTok, dN, dDst);
}

// ------------------------------------------------------------------------------------------------
// Fletcher-32 over big-endian 16-bit words                                     Lerc2.cpp:1037-1064
// With w_i (i = 0..m-1) the words of the region (odd tail byte padded with a low zero byte):
//   sum1 = 0xffff + SUM w_i ;  sum2 = 0xffff*(m+1) + SUM (m - i) * w_i      (both mod 65535, 0 -> 65535)
// (SURVEY.md Appendix B.9).  A byte at region offset r contributes c = b << (r even ? 8 : 0) to word
// r/2, so any partition of the bytes can accumulate  A = SUM c  and  D = SUM (r/2 mod 65535) * c
// independently; sum2 = 0xffff*(m+1) + m*A - D.  Each thread folds its partials mod 65535 before
// the block/global adds, so the 64-bit accumulators cannot overflow.
// 16 region bytes per thread step, big-endian 16-bit words formed with byte permutes.
// acc[0] += SUM c, acc[1] += SUM (wordIndex mod 65535) * c (mod 65535) over region bytes [0, len)
__global__ void __launch_bounds__(256) k_fletcher_partial(const uint8_t* __restrict__ region, long long len, unsigned long long* __restrict__ acc) {
  const int d = (int)((uintptr_t)region & 15);
  const uint4* g0 = (const uint4*)(region - d);
  const long long nChunks = (len + d + 15) >> 4;
  unsigned long long fa = 0, fd = 0;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nChunks; c += (long long)gridDim.x * blockDim.x) {
    uint4 x = __ldg(g0 + c);
    uint32_t o[4] = {x.x, x.y, x.z, x.w};
    const long long r0 = c * 16 - d;                               // region offset of the chunk's first byte
    if (r0 < 0 || r0 + 16 > len) {
#pragma unroll
      for (int j = 0; j < 16; j++) if (r0 + j < 0 || r0 + j >= len) o[j >> 2] &= ~(0xffu << (8 * (j & 3)));
    }
    const unsigned par = (unsigned)(r0 & 1);
    const uint32_t w0 = (uint32_t)((unsigned long long)(r0 + 16) >> 1) % 65535u;   // word index of byte r0 + 16 - par ... minus 8 below
    uint32_t S = 0, S1 = 0, prev = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const uint32_t cur = k < 4 ? o[k] : 0u;
      const uint32_t y = __funnelshift_l(prev, cur, par * 8);
      const uint32_t pw = __byte_perm(y, 0, 0x2301);
      const uint32_t wlo = pw & 0xffffu, whi = pw >> 16;
      S += wlo + whi; S1 += (uint32_t)(2 * k) * (wlo + whi) + whi;
      prev = cur;
    }
    // words of this chunk have indices w0 - 8 + k (mod 65535), k = 0..9 (index base of byte r0 - par is (r0 - par) / 2 = (r0 + 16) / 2 - 8)
    fa += S; fd += (unsigned long long)(w0 + 65535u - 8u) * S + S1;
  }
  fa %= 65535ull; fd %= 65535ull;
#pragma unroll
  for (int m = 16; m; m >>= 1) { fa += __shfl_xor_sync(FULL, fa, m); fd += __shfl_xor_sync(FULL, fd, m); }
  __shared__ unsigned long long sA[8], sD[8];
  if ((threadIdx.x & 31) == 0) { sA[threadIdx.x >> 5] = fa; sD[threadIdx.x >> 5] = fd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long A = 0, D = 0;
    for (int i = 0; i < 8; i++) { A += sA[i]; D += sD[i]; }
    if (A | D) { atomicAdd(&acc[0], A); atomicAdd(&acc[1], D % 65535ull); }
  }
}

// acc -> checksum; either stored at dst (encode) or compared with `expect` (decode: status |= 2 on mismatch)
__global__ void k_fletcher_finish(const unsigned long long* __restrict__ acc, long long len, uint8_t* dst, uint32_t expect, int* status) {
  const unsigned long long M = 65535ull;
  const unsigned long long m = (unsigned long long)((len + 1) >> 1);
  const unsigned long long A = acc[0] % M, D = acc[1] % M;
  unsigned long long s1 = (0xffffull + A) % M;
  unsigned long long s2 = ((0xffffull % M) * ((m + 1) % M) + (m % M) * A + (M - D)) % M;
  if (s1 == 0) s1 = M;
  if (s2 == 0) s2 = M;
  const uint32_t cs = (uint32_t)((s2 << 16) | s1);
  if (dst) { dst[0] = (uint8_t)cs; dst[1] = (uint8_t)(cs >> 8); dst[2] = (uint8_t)(cs >> 16); dst[3] = (uint8_t)(cs >> 24); }
  else if (cs != expect) atomicOr(status, 2);
}

void launchFletcher(Context* ctx, const uint8_t* dRegion, long long len, unsigned long long* dAcc /*2, zeroed*/,
                    uint8_t* dStoreAt, uint32_t expect, int* dStatus) {
  const long long nVec = (len + 30) >> 4;
  int grid = (int)std::min<long long>((nVec + 255) / 256, 148 * 32);
  if (grid < 1) grid = 1;
  LERC_LAUNCH(ctx, k_fletcher_partial, grid, 256, 0, dRegion, len, dAcc);
  LERC_LAUNCH(ctx, k_fletcher_finish, 1, 1, 0, dAcc, len, dStoreAt, expect, dStatus);
}

// ------------------------------------------------------------------------------------------------
// Lerc::RemapNoData (Lerc.cpp:1046-1076)
template <class T>
__global__ void k_remap_nodata(T* __restrict__ data, const uint8_t* __restrict__ bits, long long nPix, int nDepth, double fromD, double toD) {
  const T from = (T)fromD, to = (T)toD;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nPix; k += (long long)gridDim.x * blockDim.x) {
    if (!maskBit(bits, k)) continue;
    T* px = data + k * nDepth;
    for (int m = 0; m < nDepth; m++) if (px[m] == from) px[m] = to;
  }
}
void launchRemapNoData(Context* ctx, int dt, void* dData, const uint8_t* dBits, long long nPix, int nDepth, double from, double to) {
  int grid = (int)std::min<long long>((nPix + 255) / 256, 148 * 16);
  if (grid < 1) grid = 1;
  switch (dt) {
    case DT_Char:   LERC_LAUNCH(ctx, k_remap_nodata<int8_t>,   grid, 256, 0, (int8_t*)dData, dBits, nPix, nDepth, from, to); break;
    case DT_Byte:   LERC_LAUNCH(ctx, k_remap_nodata<uint8_t>,  grid, 256, 0, (uint8_t*)dData, dBits, nPix, nDepth, from, to); break;
    case DT_Short:  LERC_LAUNCH(ctx, k_remap_nodata<int16_t>,  grid, 256, 0, (int16_t*)dData, dBits, nPix, nDepth, from, to); break;
    case DT_UShort: LERC_LAUNCH(ctx, k_remap_nodata<uint16_t>, grid, 256, 0, (uint16_t*)dData, dBits, nPix, nDepth, from, to); break;
    case DT_Int:    LERC_LAUNCH(ctx, k_remap_nodata<int32_t>,  grid, 256, 0, (int32_t*)dData, dBits, nPix, nDepth, from, to); break;
    case DT_UInt:   LERC_LAUNCH(ctx, k_remap_nodata<uint32_t>, grid, 256, 0, (uint32_t*)dData, dBits, nPix, nDepth, from, to); break;
    case DT_Float:  LERC_LAUNCH(ctx, k_remap_nodata<float>,    grid, 256, 0, (float*)dData, dBits, nPix, nDepth, from, to); break;
    case DT_Double: LERC_LAUNCH(ctx, k_remap_nodata<double>,   grid, 256, 0, (double*)dData, dBits, nPix, nDepth, from, to); break;
    default: break;
  }
}

// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void k_to_double(const T* __restrict__ src, size_t n, double* __restrict__ dst) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) dst[k] = (double)src[k];
}
void launchConvertToDouble(Context* ctx, const void* dSrc, int dt, size_t n, double* dDst) {
  int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 32);
  if (grid < 1) grid = 1;
  switch (dt) {
    case DT_Char:   LERC_LAUNCH(ctx, k_to_double<int8_t>,   grid, 256, 0, (const int8_t*)dSrc, n, dDst); break;
    case DT_Byte:   LERC_LAUNCH(ctx, k_to_double<uint8_t>,  grid, 256, 0, (const uint8_t*)dSrc, n, dDst); break;
    case DT_Short:  LERC_LAUNCH(ctx, k_to_double<int16_t>,  grid, 256, 0, (const int16_t*)dSrc, n, dDst); break;
    case DT_UShort: LERC_LAUNCH(ctx, k_to_double<uint16_t>, grid, 256, 0, (const uint16_t*)dSrc, n, dDst); break;
    case DT_Int:    LERC_LAUNCH(ctx, k_to_double<int32_t>,  grid, 256, 0, (const int32_t*)dSrc, n, dDst); break;
    case DT_UInt:   LERC_LAUNCH(ctx, k_to_double<uint32_t>, grid, 256, 0, (const uint32_t*)dSrc, n, dDst); break;
    case DT_Float:  LERC_LAUNCH(ctx, k_to_double<float>,    grid, 256, 0, (const float*)dSrc, n, dDst); break;
    default: break;
  }
}

}  // namespace lerc
