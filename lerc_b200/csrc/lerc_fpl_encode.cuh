// lerc_fpl_encode.cuh -- lossless float / double codec ("FPL", IEM_DeltaDeltaHuffman), ENCODE side.  Included by lerc_encode.cu.
//
// What the reference does serially (fpl_Lerc2Ext.cpp:438-608, fpl_Compression.cpp:53-112, fpl_EsriHuffman.cpp:82-452,
// fpl_UnitTypes.cpp:39-136, :302-517) is split here into data-parallel kernels plus a few host decisions on histograms:
//
//   k_fpl_transform     float bits -> exponent | sign | mantissa (moveBits2Front); NaN at a valid pixel -> 0 (Lerc.cpp:1437-1441)
//   k_fpl_test_hist     the three predictors (none, row delta, row + column delta) are evaluated on the fly on the reference's
//                       test blocks: per block, byte plane and "first byte delta" variant a histogram of every 7th byte
//   host                entropy estimates in double (log2 of the host libm, like the reference) -> predictor
//   k_fpl_planes        predictor applied, value bytes split into planes
//   k_fpl_level_hist    per plane and 8 KB snippet: histograms of every 7th byte for byte-delta levels 0..maxDelta
//   host                best level per plane (stops at the first level that does not improve, fpl_Lerc2Ext.cpp:303-318)
//   k_fpl_derive        the plane's level-th order byte differences (order min(i, level) at index i, as the in-place loops give)
//   k_fpl_hist          full histogram per plane -> host: canonical Huffman table and its coded size
//   k_pb_*              PackBits size of the plane from three scans (run starts, literal-sequence starts, output offsets)
//   host                one value / PackBits / stored / Huffman, like fpl_EsriHuffman.cpp:316-452
//   k_pb_write | copy | k_fpl_huff_bits + k_fpl_huff_write      the plane payload, kept in scratch until the band is written
//
// One deviation, on purpose: the reference leaves the read-ahead word behind a Huffman-coded plane uninitialised (malloc'ed,
// never written: fpl_EsriHuffman.cpp:403-448); zeros are written here, as the 8-bit Huffman path of Lerc2 does.
#pragma once

namespace lerc {
namespace {

constexpr int FPL_PRIME = 7, FPL_MAX_DELTA = 5, FPL_SAMPLE = 8 * 1024;

template <class U> __device__ __forceinline__ U fplSub(U a, U b);
template <> __device__ __forceinline__ uint32_t fplSub<uint32_t>(uint32_t a, uint32_t b) {                       // fpl_UnitTypes.cpp:83-97
  return ((a - b) & 0x007FFFFFu) | (((((a >> 23) & 0x1FFu) - ((b >> 23) & 0x1FFu)) & 0x1FFu) << 23);
}
template <> __device__ __forceinline__ unsigned long long fplSub<unsigned long long>(unsigned long long a, unsigned long long b) {   // :119-136
  const unsigned long long M = 0x000FFFFFFFFFFFFFull;
  return (((a & M) - (b & M)) & M) | (((((a >> 52) & 0xFFFull) - ((b >> 52) & 0xFFFull)) & 0xFFFull) << 52);
}

// element (r, c) after predictor 0 (none), 1 (difference to the left neighbour) or 2 (that, then difference to the row above);
// the reference's in-place loops run backwards, so every difference sees its neighbour's old value (fpl_UnitTypes.cpp:302-357, :436-517)
template <class U>
__device__ __forceinline__ U fplPredicted(const U* __restrict__ X, unsigned long long cols, unsigned long long r, unsigned long long c, int pred) {
  const U* row = X + r * cols;
  const U v = row[c];
  if (pred == 0) return v;
  const U d = c ? fplSub<U>(v, row[c - 1]) : v;
  if (pred == 1 || r == 0) return d;
  const U* up = row - cols;
  const U du = c ? fplSub<U>(up[c], up[c - 1]) : up[c];
  return fplSub<U>(d, du);
}

template <class T, class U>
__global__ void k_fpl_transform(const T* __restrict__ data, const uint8_t* __restrict__ validBytes, unsigned long long n, int nanToZero,
                                U* __restrict__ X) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    U u;
    if constexpr (sizeof(T) == 4) {
      u = __float_as_uint(data[i]);
      if (nanToZero && (u & 0x7fffffffu) > 0x7f800000u && (!validBytes || validBytes[i])) u = 0;
      u = (u & 0x007FFFFFu) | (((u >> 23) & 0xFFu) << 24) | ((u >> 31) << 23);                                    // fpl_UnitTypes.cpp:39-51
    } else {
      u = (U)__double_as_longlong(data[i]);
      if (nanToZero && (u & 0x7fffffffffffffffull) > 0x7ff0000000000000ull && (!validBytes || validBytes[i])) u = 0;
    }
    X[i] = u;
  }
}

struct FplBlock { long long top, height; };

// grid (nBlk, 3 predictors).  out[(((pred * nBlk + blk) * unit + byte) * 2 + variant) * 256 + value]; variant 1 = every sampled
// byte but the first minus its predecessor in the plane (setDerivativePrime, fpl_Lerc2Ext.cpp:103-116)
template <class U>
__global__ void k_fpl_test_hist(const U* __restrict__ X, unsigned long long cols, const FplBlock* __restrict__ blocks, int* __restrict__ out) {
  constexpr int unit = (int)sizeof(U);
  __shared__ int h[unit * 2 * 256];
  for (int i = threadIdx.x; i < unit * 2 * 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const FplBlock blk = blocks[blockIdx.x];
  const int pred = (int)blockIdx.y;
  const unsigned long long length = (unsigned long long)blk.height * cols, first = (unsigned long long)blk.top * cols;
  const unsigned long long nSamp = (length + FPL_PRIME - 1) / FPL_PRIME;
  for (unsigned long long s = threadIdx.x; s < nSamp; s += blockDim.x) {
    const unsigned long long i = s * FPL_PRIME, e = first + i;
    const U v = fplPredicted<U>(X, cols, e / cols, e % cols, pred);
    const U w = i ? fplPredicted<U>(X, cols, (e - 1) / cols, (e - 1) % cols, pred) : (U)0;
#pragma unroll
    for (int byte = 0; byte < unit; byte++) {
      const unsigned b = (unsigned)(v >> (8 * byte)) & 255u, pb = (unsigned)(w >> (8 * byte)) & 255u;
      atomicAdd(&h[(byte * 2 + 0) * 256 + b], 1);
      atomicAdd(&h[(byte * 2 + 1) * 256 + (i ? ((b - pb) & 255u) : b)], 1);
    }
  }
  __syncthreads();
  int* dst = out + ((size_t)pred * gridDim.x + blockIdx.x) * (size_t)(unit * 2 * 256);
  for (int i = threadIdx.x; i < unit * 2 * 256; i += blockDim.x) dst[i] = h[i];
}

template <class U>
__global__ void k_fpl_planes(const U* __restrict__ X, unsigned long long cols, unsigned long long n, int pred, uint8_t* __restrict__ P) {
  constexpr int unit = (int)sizeof(U);
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x) {
    const U v = fplPredicted<U>(X, cols, e / cols, e % cols, pred);
#pragma unroll
    for (int byte = 0; byte < unit; byte++) P[(size_t)byte * n + e] = (uint8_t)(v >> (8 * byte));
  }
}

struct FplSnippet { unsigned int start, len; };

// grid (nSnip, unit planes).  out[((plane * nSnip + s) * (FPL_MAX_DELTA + 1) + level) * 256 + value]      fpl_Lerc2Ext.cpp:283-318
__global__ void k_fpl_level_hist(const uint8_t* __restrict__ P, unsigned long long n, const FplSnippet* __restrict__ snips, int maxDelta,
                                 int* __restrict__ out) {
  __shared__ uint8_t bufA[FPL_SAMPLE], bufB[FPL_SAMPLE];
  __shared__ int h[256];
  const FplSnippet sn = snips[blockIdx.x];
  const uint8_t* src = P + (size_t)blockIdx.y * n + sn.start;
  uint8_t* cur = bufA; uint8_t* nxt = bufB;
  for (unsigned i = threadIdx.x; i < sn.len; i += blockDim.x) cur[i] = src[i];
  for (int l = 0; l <= maxDelta; l++) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    if (l > 0) {
      for (unsigned i = threadIdx.x; i < sn.len; i += blockDim.x) nxt[i] = i >= (unsigned)l ? (uint8_t)(cur[i] - cur[i - 1]) : cur[i];
      __syncthreads();
      uint8_t* t = cur; cur = nxt; nxt = t;
    }
    for (unsigned i = threadIdx.x * FPL_PRIME; i < sn.len; i += blockDim.x * FPL_PRIME) atomicAdd(&h[cur[i]], 1);
    __syncthreads();
    int* dst = out + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (FPL_MAX_DELTA + 1) + l) * 256;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) dst[i] = h[i];
    __syncthreads();
  }
}

// setDerivative (fpl_Lerc2Ext.cpp:118-131): pass l = 1..level subtracts the left neighbour at every index >= l
__global__ void k_fpl_derive(const uint8_t* __restrict__ in, unsigned long long n, int level, uint8_t* __restrict__ out) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const int m = i < (unsigned long long)level ? (int)i : level;
    uint8_t t[FPL_MAX_DELTA + 1];
    for (int k = 0; k <= m; k++) t[k] = in[i - (unsigned long long)(m - k)];
    for (int l = 1; l <= m; l++)
      for (int k = m; k >= l; k--) t[k] = (uint8_t)(t[k] - t[k - 1]);
    out[i] = t[m];
  }
}

// grid (x, planes): hist[plane * 256 + value]
__global__ void k_fpl_hist(const uint8_t* __restrict__ P, unsigned long long n, int* __restrict__ hist) {
  __shared__ int h[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const uint8_t* p = P + (size_t)blockIdx.y * n;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) atomicAdd(&h[p[i]], 1);
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) if (h[i]) atomicAdd(&hist[blockIdx.y * 256 + i], h[i]);
}

// ---- PackBits (fpl_EsriHuffman.cpp:82-166).  A maximal run of equal bytes is cut into chunks of 129 from its start; a chunk of
// 2..129 bytes is a repeat token {127 + (len - 1), byte}, a chunk of one byte is a literal; adjacent literals form groups of up
// to 128, each behind a count byte {n - 1}.
struct FplMaxOp { __host__ __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };

__global__ void k_pb_run_flags(const uint8_t* __restrict__ p, unsigned long long n, uint32_t* __restrict__ f) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
    f[i] = (i > 0 && p[i] != p[i - 1]) ? (uint32_t)i : 0u;
}
__device__ __forceinline__ bool pbChunkStart(const uint32_t* runStart, unsigned long long i) { return (i - runStart[i]) % 129u == 0; }
__device__ __forceinline__ bool pbIsLiteral(const uint8_t* p, const uint32_t* runStart, unsigned long long n, unsigned long long i) {
  return pbChunkStart(runStart, i) && (i + 1 == n || p[i + 1] != p[i]);
}
// g[i] = i + 1 where a literal sequence starts, else 0
__global__ void k_pb_lit_flags(const uint8_t* __restrict__ p, const uint32_t* __restrict__ runStart, unsigned long long n, uint32_t* __restrict__ g) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
    g[i] = (pbIsLiteral(p, runStart, n, i) && !(i > 0 && pbIsLiteral(p, runStart, n, i - 1))) ? (uint32_t)(i + 1) : 0u;
}
// bytes each input position contributes; c[n] = 0 so that the exclusive sum over n + 1 entries ends with the total
__global__ void k_pb_contrib(const uint8_t* __restrict__ p, const uint32_t* __restrict__ runStart, const uint32_t* __restrict__ litStart1,
                             unsigned long long n, uint32_t* __restrict__ c) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (unsigned long long)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    if (i < n && pbChunkStart(runStart, i)) {
      if (pbIsLiteral(p, runStart, n, i)) v = 1u + (((i - (litStart1[i] - 1u)) % 128u) == 0 ? 1u : 0u);
      else v = 2;
    }
    c[i] = v;
  }
}
__global__ void k_pb_write(const uint8_t* __restrict__ p, const uint32_t* __restrict__ runStart, const uint32_t* __restrict__ litStart1,
                           const uint32_t* __restrict__ off, unsigned long long n, uint8_t* __restrict__ out) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    if (!pbChunkStart(runStart, i)) continue;
    const uint8_t b = p[i];
    uint32_t o = off[i];
    if (pbIsLiteral(p, runStart, n, i)) {
      if (((i - (litStart1[i] - 1u)) % 128u) == 0) {                          // first of its group: count the group
        int cnt = 1;
        while (cnt < 128 && i + cnt < n && pbIsLiteral(p, runStart, n, i + cnt)) cnt++;
        out[o++] = (uint8_t)(cnt - 1);
      }
      out[o] = b;
    } else {
      int rep = 1;
      while (rep < 128 && i + rep + 1 < n && p[i + rep + 1] == b) rep++;
      out[o] = (uint8_t)(127 + rep); out[o + 1] = b;
    }
  }
}

// ---- Huffman bit stream of one plane (fpl_EsriHuffman.cpp:411-441): MSB first in little-endian uint32 words.  One warp per
// segment of 1024 symbols, one lane per 32 consecutive symbols.
struct FplHuffArgs { const uint8_t* in; unsigned long long n; uint16_t len[256]; uint32_t code[256]; };

__global__ void k_fpl_huff_bits(FplHuffArgs a, unsigned long long nSeg, unsigned long long* __restrict__ segBits) {
  const int lane = threadIdx.x & 31;
  for (unsigned long long seg = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); seg < nSeg; seg += (unsigned long long)gridDim.x * (blockDim.x >> 5)) {
    const unsigned long long i0 = seg * 1024ull + (unsigned long long)lane * 32ull;
    unsigned bits = 0;
    for (int k = 0; k < 32; k++) if (i0 + k < a.n) bits += a.len[a.in[i0 + k]];
    bits = __reduce_add_sync(FULL, bits);
    if (lane == 0) segBits[seg] = bits;
  }
}
__global__ void k_fpl_huff_write(FplHuffArgs a, unsigned long long nSeg, const unsigned long long* __restrict__ segOff, uint32_t* __restrict__ words) {
  const int lane = threadIdx.x & 31;
  for (unsigned long long seg = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); seg < nSeg; seg += (unsigned long long)gridDim.x * (blockDim.x >> 5)) {
    const unsigned long long i0 = seg * 1024ull + (unsigned long long)lane * 32ull;
    unsigned bits = 0;
    for (int k = 0; k < 32; k++) if (i0 + k < a.n) bits += a.len[a.in[i0 + k]];
    unsigned incl = bits;
    for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
    unsigned long long pos = segOff[seg] + (incl - bits);
    for (int k = 0; k < 32; k++) {
      if (i0 + k >= a.n) break;
      const unsigned s = a.in[i0 + k];
      const int len = a.len[s];
      const uint32_t code = a.code[s];
      const unsigned long long w = pos >> 5;
      const int sh = (int)(pos & 31);
      if (32 - sh >= len) atomicOr(&words[w], code << (32 - sh - len));
      else { const int rem = len - (32 - sh); atomicOr(&words[w], code >> rem); atomicOr(&words[w + 1], code << (32 - rem)); }
      pos += (unsigned)len;
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
struct FplPlanePlan {
  int level = 0;
  uint32_t size = 0;                  // coded bytes of the plane, flag byte included
  std::vector<uint8_t> head;          // written by the host: flag byte [+ value and count | Huffman table]
  const uint8_t* dPayload = nullptr;  // written from scratch memory behind `head`
  size_t payloadBytes = 0;
};
struct FplPlan {
  int pred = 0, nPlanes = 0;
  FplPlanePlan pl[8];
  long long bytes() const { long long r = 1; for (int i = 0; i < nPlanes; i++) r += (long long)pl[i].size + 6; return r; }   // fpl_Lerc2Ext.cpp:397-408
};

// fpl_Compression.cpp:85-112 on a histogram of the sampled bytes
inline long fplEntropyBytes(const int* table) {
  long long total = 0;
  for (int i = 0; i < 256; i++) total += table[i];
  const int totalCount = (int)total;
  double bitsSum = 0;
  for (int i = 0; i < 256; i++) {
    if (table[i] == 0) continue;
    const double p = (double)totalCount / (unsigned long)table[i];
    const double bits = std::log2(p);
    bitsSum += bits * (unsigned long)table[i];
  }
  return (long)((bitsSum + 7) / 8);
}

// fpl_Lerc2Ext.cpp:62-101
inline std::vector<FplBlock> fplTestBlocks(int width, int height) {
  const size_t size = (size_t)width * (size_t)height;
  const double t = std::round((double)size / FPL_SAMPLE);
  int count = (int)std::round(std::sqrt(t + 1));
  int bh = FPL_SAMPLE / width;
  if (bh < 4) bh = 4;
  while (count * bh > height && count > 1) count--;
  const float topMargin = (float)((height - count * bh) / (2.0 * count));
  const float delta = 2.0f * topMargin + bh;
  std::vector<FplBlock> v;
  for (int i = 0; i < count; i++) {
    FplBlock tb; tb.top = (long long)(topMargin + delta * i); tb.height = bh;
    if (tb.top < 0) tb.top = 0;
    if (tb.top + tb.height > height) tb.height = height - tb.top;
    if (tb.height > 0) v.push_back(tb);
  }
  return v;
}

// fpl_Lerc2Ext.cpp:238-279
inline std::vector<FplSnippet> fplSnippets(size_t size) {
  const unsigned target = FPL_SAMPLE;
  const double t = std::round((double)size / target);
  int count = (int)std::round(std::sqrt(t + 1));
  while ((size_t)((unsigned)count * target) > size && count > 0) count--;
  std::vector<FplSnippet> v;
  if (count == 0) return v;
  const float topMargin = (float)(((unsigned)(int)size - (unsigned)count * target) / (2.0 * count));
  const float delta = 2.0f * topMargin + target;
  for (int i = 0; i < count; i++) {
    long st = (long)(topMargin + delta * i); int ln = (int)target;
    if (st < 0) st = 0;
    if (st + ln > (int)size) ln = (int)size - (int)st;
    if (ln > 0) v.push_back(FplSnippet{(unsigned)st, (unsigned)ln});
  }
  return v;
}

inline bool fplMaxScan(Context* ctx, uint32_t* d, size_t n) {
  LaunchScope scope(ctx, "cub::InclusiveScan<max>");
  size_t tmpBytes = 0;
  cub::DeviceScan::InclusiveScan(nullptr, tmpBytes, d, d, FplMaxOp(), (int)n, ctx->stream);
  void* tmp = ctx->arena.alloc(tmpBytes ? tmpBytes : 16);
  if (!tmp) return false;
  cub::DeviceScan::InclusiveScan(tmp, tmpBytes, d, d, FplMaxOp(), (int)n, ctx->stream);
  ctx->kernelLaunches += 2;
  return true;
}

template <class V> bool d2h(Context* ctx, V* hostDst, const void* dSrc, size_t count);

// Sizing = coding: every decision of ComputeHuffmanCodesFltSlice (fpl_Lerc2Ext.cpp:455-608) and the coded planes, left in scratch.
template <class T>
bool planFpl(Context* ctx, const T* dData, const uint8_t* dValidBytes, int nCols, int nRows, int nDepth, FplPlan& plan) {
  using U = typename std::conditional<sizeof(T) == 4, uint32_t, unsigned long long>::type;
  constexpr int unit = (int)sizeof(T);
  cudaStream_t st = ctx->stream;
  const unsigned long long cols = nDepth == 1 ? (unsigned long long)nCols : (unsigned long long)nDepth;       // fpl_Lerc2Ext.cpp:438-453
  const unsigned long long rows = nDepth == 1 ? (unsigned long long)nRows : (unsigned long long)nCols * (unsigned long long)nRows;
  const unsigned long long n = cols * rows;
  if (n == 0 || n > 0x7fffffffull) return false;
  const int fillGrid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((n + 255) / 256, 148ull * 32));

  U* X = (U*)ctx->arena.alloc((size_t)n * sizeof(U));
  uint8_t* P = (uint8_t*)ctx->arena.alloc((size_t)n * unit + 16);
  uint8_t* Q = (uint8_t*)ctx->arena.alloc((size_t)n * unit + 16);                                              // planes after the byte deltas
  if (!X || !P || !Q) return false;
  LERC_LAUNCH(ctx, (k_fpl_transform<T, U>), fillGrid, 256, 0, dData, dValidBytes, n, nDepth == 1 ? 1 : 0, X);

  // ---- predictor (selectInitialLinearOrCrossDelta :341-395 on the test blocks)
  const std::vector<FplBlock> blocks = fplTestBlocks((int)cols, (int)rows);
  int pred = 0;
  if (!blocks.empty()) {
    const size_t nBlk = blocks.size(), histInts = 3 * nBlk * (size_t)unit * 2 * 256;
    FplBlock* dBlocks = (FplBlock*)ctx->arena.alloc(nBlk * sizeof(FplBlock));
    int* dHist = (int*)ctx->arena.alloc(histInts * 4);
    if (!dBlocks || !dHist) return false;
    if (!cudaOk(cudaMemcpyAsync(dBlocks, blocks.data(), nBlk * sizeof(FplBlock), cudaMemcpyHostToDevice, st), "H2D fpl blocks")) return false;
    if (!cudaOk(cudaStreamSynchronize(st), "sync")) return false;                                              // `blocks` is pageable
    LERC_LAUNCH(ctx, k_fpl_test_hist<U>, dim3((unsigned)nBlk, 3), 256, 0, (const U*)X, cols, (const FplBlock*)dBlocks, dHist);
    std::vector<int> hist(histInts);
    if (!d2h(ctx, hist.data(), dHist, histInts)) return false;
    size_t stats[3] = {0, 0, 0};
    for (int p = 0; p < 3; p++)
      for (size_t b = 0; b < nBlk; b++)
        for (int byte = 0; byte < unit; byte++) {
          const int* h0 = hist.data() + (((size_t)p * nBlk + b) * unit + byte) * 2 * 256;
          const size_t e1 = (size_t)fplEntropyBytes(h0), e2 = (size_t)fplEntropyBytes(h0 + 256);
          stats[p] += std::min(e1, e2);
        }
    if (stats[1] < stats[pred]) pred = 1;
    if (stats[2] < stats[pred]) pred = 2;
  }
  plan.pred = pred;
  LERC_LAUNCH(ctx, k_fpl_planes<U>, fillGrid, 256, 0, (const U*)X, cols, n, pred, P);

  // ---- byte-delta level per plane (getBestLevel2 :238-330)
  const int maxDelta = FPL_MAX_DELTA - pred;                                                                    // fpl_Predictor.cpp:30-33
  int level[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const std::vector<FplSnippet> snips = fplSnippets((size_t)n);
  if (!snips.empty() && maxDelta > 0) {
    const size_t nSnip = snips.size(), histInts = (size_t)unit * nSnip * (FPL_MAX_DELTA + 1) * 256;
    FplSnippet* dSnips = (FplSnippet*)ctx->arena.alloc(nSnip * sizeof(FplSnippet));
    int* dHist = (int*)ctx->arena.alloc(histInts * 4);
    if (!dSnips || !dHist) return false;
    if (!cudaOk(cudaMemcpyAsync(dSnips, snips.data(), nSnip * sizeof(FplSnippet), cudaMemcpyHostToDevice, st), "H2D fpl snippets")) return false;
    if (!cudaOk(cudaStreamSynchronize(st), "sync")) return false;
    cudaMemsetAsync(dHist, 0, histInts * 4, st);
    LERC_LAUNCH(ctx, k_fpl_level_hist, dim3((unsigned)nSnip, (unsigned)unit), 256, 0, (const uint8_t*)P, n, (const FplSnippet*)dSnips, maxDelta, dHist);
    std::vector<int> hist(histInts);
    if (!d2h(ctx, hist.data(), dHist, histInts)) return false;
    for (int byte = 0; byte < unit; byte++) {
      size_t best = 0;
      for (int l = 0; l <= maxDelta; l++) {
        size_t comp = 0;
        for (size_t s = 0; s < nSnip; s++) comp += (size_t)fplEntropyBytes(hist.data() + (((size_t)byte * nSnip + s) * (FPL_MAX_DELTA + 1) + l) * 256);
        if (comp < best || l == 0) { best = comp; level[byte] = l; } else break;
      }
    }
  }
  for (int byte = 0; byte < unit; byte++) {
    if (level[byte] == 0) cudaMemcpyAsync(Q + (size_t)byte * n, P + (size_t)byte * n, (size_t)n, cudaMemcpyDeviceToDevice, st);
    else LERC_LAUNCH(ctx, k_fpl_derive, fillGrid, 256, 0, (const uint8_t*)(P + (size_t)byte * n), n, level[byte], Q + (size_t)byte * n);
  }

  // ---- plane coding (fpl_EsriHuffman.cpp:316-452)
  int* dPlaneHist = (int*)ctx->arena.alloc((size_t)unit * 256 * 4);
  if (!dPlaneHist) return false;
  cudaMemsetAsync(dPlaneHist, 0, (size_t)unit * 256 * 4, st);
  LERC_LAUNCH(ctx, k_fpl_hist, dim3((unsigned)std::min<unsigned long long>((n + 255) / 256, 148ull * 8), (unsigned)unit), 256, 0, (const uint8_t*)Q, n, dPlaneHist);
  int planeHist[8 * 256];
  if (!d2h(ctx, planeHist, dPlaneHist, (size_t)unit * 256)) return false;

  // PackBits scratch, shared by the planes (stream order keeps the uses apart)
  uint32_t* dRun = (uint32_t*)ctx->arena.alloc(((size_t)n + 1) * 4);
  uint32_t* dLit = (uint32_t*)ctx->arena.alloc(((size_t)n + 1) * 4);
  uint32_t* dCon = (uint32_t*)ctx->arena.alloc(((size_t)n + 2) * 4);
  uint32_t* dOff = (uint32_t*)ctx->arena.alloc(((size_t)n + 2) * 4);
  if (!dRun || !dLit || !dCon || !dOff) return false;

  plan.nPlanes = 0;
  for (int byte = 0; byte < unit; byte++) {
    FplPlanePlan& pp = plan.pl[byte];
    pp = FplPlanePlan();
    pp.level = level[byte];
    const uint8_t* q = Q + (size_t)byte * n;
    const int* histo = planeHist + byte * 256;
    int distinct = 0, firstVal = 0;
    for (int i = 255; i >= 0; i--) if (histo[i] > 0) { distinct++; firstVal = i; }
    if (distinct < 2) {                                                                                         // :323-343: flag, value, count
      const uint32_t cnt = (uint32_t)n;
      pp.head.assign(6, 0); pp.head[0] = 1; pp.head[1] = (uint8_t)firstVal; std::memcpy(pp.head.data() + 2, &cnt, 4);
      pp.size = 6;
    } else {
      HuffmanTable ht; int numBytes = 0;
      if (!ht.buildFromHistogram(histo) || !ht.totalBytes(histo, numBytes) || numBytes <= 0) return false;       // :345-348
      LERC_LAUNCH(ctx, k_pb_run_flags, fillGrid, 256, 0, q, n, dRun);
      if (!fplMaxScan(ctx, dRun, (size_t)n)) return false;
      LERC_LAUNCH(ctx, k_pb_lit_flags, fillGrid, 256, 0, q, (const uint32_t*)dRun, n, dLit);
      if (!fplMaxScan(ctx, dLit, (size_t)n)) return false;
      LERC_LAUNCH(ctx, k_pb_contrib, fillGrid, 256, 0, q, (const uint32_t*)dRun, (const uint32_t*)dLit, n, dCon);
      exclusiveScanU32(ctx, dCon, dOff, (size_t)n);
      uint32_t pb = 0;
      if (!d2h(ctx, &pb, dOff + n, 1)) return false;
      if (pb > 0 && (long long)pb < (long long)numBytes && (unsigned long long)pb < n) {                        // :350-372
        uint8_t* dPb = (uint8_t*)ctx->arena.alloc((size_t)pb + 16);
        if (!dPb) return false;
        LERC_LAUNCH(ctx, k_pb_write, fillGrid, 256, 0, q, (const uint32_t*)dRun, (const uint32_t*)dLit, (const uint32_t*)dOff, n, dPb);
        pp.head.assign(1, 3); pp.dPayload = dPb; pp.payloadBytes = pb; pp.size = pb + 1;
      } else if ((unsigned long long)numBytes >= n) {                                                           // :374-387: stored
        pp.head.assign(1, 2); pp.dPayload = q; pp.payloadBytes = (size_t)n; pp.size = (uint32_t)(n + 1);
      } else {                                                                                                  // :389-451: Huffman
        std::vector<uint8_t> tb(4096, 0);
        const size_t tableBytes = ht.write(tb.data(), 5);
        if (!tableBytes || tableBytes >= (size_t)numBytes) return false;
        const size_t dataBytes = (size_t)numBytes - tableBytes;                                                  // bit stream words + the read-ahead word
        uint32_t* dWords = (uint32_t*)ctx->arena.alloc(dataBytes + 16);
        const unsigned long long nSeg = (n + 1023) / 1024;
        unsigned long long* dSegBits = (unsigned long long*)ctx->arena.alloc(8 * ((size_t)nSeg + 1));
        unsigned long long* dSegOff = (unsigned long long*)ctx->arena.alloc(8 * ((size_t)nSeg + 1));
        if (!dWords || !dSegBits || !dSegOff) return false;
        cudaMemsetAsync(dWords, 0, dataBytes + 16, st);
        cudaMemsetAsync(dSegBits + nSeg, 0, 8, st);
        FplHuffArgs ha; ha.in = q; ha.n = n;
        std::memcpy(ha.len, ht.len, sizeof ha.len); std::memcpy(ha.code, ht.code, sizeof ha.code);
        const int grid = (int)std::min<unsigned long long>((nSeg + 7) / 8, 148ull * 16);
        LERC_LAUNCH(ctx, k_fpl_huff_bits, grid, 256, 0, ha, nSeg, dSegBits);
        exclusiveScanU64(ctx, dSegBits, dSegOff, (size_t)nSeg);
        LERC_LAUNCH(ctx, k_fpl_huff_write, grid, 256, 0, ha, nSeg, (const unsigned long long*)dSegOff, dWords);
        pp.head.assign(1 + tableBytes, 0); std::memcpy(pp.head.data() + 1, tb.data(), tableBytes);
        pp.dPayload = (const uint8_t*)dWords; pp.payloadBytes = dataBytes; pp.size = (uint32_t)(1 + numBytes);
      }
    }
    plan.nPlanes = byte + 1;
  }
  return cudaOk(cudaGetLastError(), "planFpl");
}

// EncodeHuffmanFlt (fpl_Lerc2Ext.cpp:410-436): predictor code, then per plane {index, level, coded size, coded bytes}
inline bool writeFpl(Context* ctx, const FplPlan& plan, uint8_t* dDst, size_t& written) {
  cudaStream_t st = ctx->stream;
  size_t hostBytes = 1;
  for (int i = 0; i < plan.nPlanes; i++) hostBytes += 6 + plan.pl[i].head.size();
  uint8_t* stage = (uint8_t*)ctx->pinnedAlloc(hostBytes);
  if (!stage) return false;
  size_t sp = 0, pos = 0;
  stage[sp] = (uint8_t)plan.pred;
  cudaMemcpyAsync(dDst + pos, stage + sp, 1, cudaMemcpyHostToDevice, st);
  sp += 1; pos += 1;
  for (int i = 0; i < plan.nPlanes; i++) {
    const FplPlanePlan& pp = plan.pl[i];
    uint8_t* h = stage + sp;
    h[0] = (uint8_t)i; h[1] = (uint8_t)pp.level; std::memcpy(h + 2, &pp.size, 4);
    std::memcpy(h + 6, pp.head.data(), pp.head.size());
    const size_t hb = 6 + pp.head.size();
    cudaMemcpyAsync(dDst + pos, h, hb, cudaMemcpyHostToDevice, st);
    sp += hb; pos += hb;
    if (pp.payloadBytes) { cudaMemcpyAsync(dDst + pos, pp.dPayload, pp.payloadBytes, cudaMemcpyDeviceToDevice, st); pos += pp.payloadBytes; }
    if (pp.head.size() + pp.payloadBytes != (size_t)pp.size) return false;
  }
  written = pos;
  return true;
}

}  // namespace
}  // namespace lerc
