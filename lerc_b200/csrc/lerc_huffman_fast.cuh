// lerc_huffman_fast.cuh -- parallel decoder of the 8-bit Huffman image modes (IEM_Huffman / IEM_DeltaHuffman) for
// rasters where every pixel is valid (included by lerc_decode.cu).
//
// The reference decodes ONE bit stream for the whole band serially (Lerc2::DecodeHuffman, Lerc2.cpp:2472-2606;
// Huffman::DecodeOneValue, Huffman.h:144-214): codes are MSB-first in little-endian 32-bit words, at most 32 bits long,
// canonical in the sense of Huffman.cpp:541-572 (codes of one length are consecutive integers).  Here:
//   k_huff_chunks   the stream is cut into chunks of HF_CHUNK bits, one thread each.  Thread t decodes from its current
//                   start bit to the end of its chunk and records where it stopped and how many symbols it saw.
//                   Iteration 0 starts at the chunk boundary (usually mid-code); iteration i+1 starts where chunk t-1
//                   stopped in iteration i.  Huffman codes self-synchronise within a few symbols, so after 2-4 iterations
//                   nothing changes any more; then start[t] == end[t-1] for every t with start[0] = 0, i.e. the chain IS
//                   the serial decode (exact, not heuristic).  Only chunks whose start moved are decoded again.
//   (CUB scan)      symbol index of every chunk's first symbol
//   k_huff_emit     every thread decodes its chunk once more and stores the symbols: straight into the raster for
//                   IEM_Huffman (pixel-interleaved order, Lerc2.cpp:2590-2600), into a depth-planar delta image for
//                   IEM_DeltaHuffman (Lerc2.cpp:2501-2522)
//   k_huff_col0, k_huff_rows   undo the predictor of the delta image when all pixels are valid: column 0 is a running sum
//                   down the rows (predictor = pixel above), every row a running sum along the row (predictor = left
//                   neighbour), modulo 256, per depth plane.
// Anything unexpected (no convergence, a bit pattern that is no code on the true chain, a short stream) makes the host
// run the serial kernel (k_huffman_decode_seq), which decides about errors like the reference.
#pragma once

namespace lerc {

constexpr int HF_CHUNK = 2048;         // bits per chunk

struct HuffFastTables {                // device copy of the code book in decode form
  uint16_t lut[4096];                  // 12 leading bits -> (len << 8) | symbol for codes of <= 12 bits, 0 = longer code
  uint32_t first[33];                  // per length: smallest code
  uint32_t count[33];                  // per length: number of codes
  uint32_t offset[33];                 // per length: index of its first symbol in `syms`
  uint8_t syms[256];                   // symbols sorted by (length, code)
  int minLen, maxLen;
};

struct HuffFastArgs {
  const uint8_t* stream; unsigned long long nBits;     // bit stream and the number of bits that may be read (whole words)
  const HuffFastTables* tab;
  unsigned long long nSym;                              // symbols the band holds
  int nChunks;
  unsigned long long* startA; unsigned long long* endA; unsigned long long* endB;   // [nChunks]
  uint32_t* count;                                      // [nChunks + 1]
  int* changed; int* bad;
  unsigned long long* endBit;                           // bit position after the band's last symbol
};

// 32 bits starting at bit position p (MSB first inside little-endian words), stream pointer of any alignment
__device__ __forceinline__ uint32_t hfPeek(const uint8_t* __restrict__ s, unsigned long long p) {
  const unsigned long long byte = (p >> 5) * 4;
  const uintptr_t ad = (uintptr_t)(s + byte);
  const uint32_t* w = (const uint32_t*)(ad & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(ad & 3) * 8;
  const uint32_t a = __ldg(w), b = __ldg(w + 1), c = __ldg(w + 2);
  const uint32_t w0 = __funnelshift_r(a, b, sh), w1 = __funnelshift_r(b, c, sh);   // the two stream words
  const uint32_t o = (uint32_t)(p & 31);
  return __funnelshift_l(w1, w0, o);                                                // (w0 << o) | (w1 >> (32 - o))
}

// Sequential MSB-first bit reader over little-endian 32-bit stream words (stream pointer of any alignment):
// 64-bit window, one aligned 32-bit load per 32 consumed bits.
struct HfReader {
  const uint32_t* al; uint32_t sh;      // aligned word pointer of stream word `next`, byte misalignment * 8
  uint32_t carry;                       // aligned word already loaded (low part of the next stream word)
  unsigned long long buf; int avail;    // bits of the stream starting at the current position, left aligned
  __device__ __forceinline__ uint32_t word() {                     // next stream word
    const uint32_t hi = __ldg(al + 1);
    const uint32_t w = __funnelshift_r(carry, hi, sh);
    carry = hi; al++;
    return w;
  }
  __device__ __forceinline__ void init(const uint8_t* __restrict__ s, unsigned long long p) {
    const uintptr_t ad = (uintptr_t)(s + (p >> 5) * 4);
    al = (const uint32_t*)(ad & ~(uintptr_t)3); sh = (uint32_t)(ad & 3) * 8;
    carry = __ldg(al);
    const uint32_t w0 = word(), w1 = word();
    const uint32_t o = (uint32_t)(p & 31);
    buf = (((unsigned long long)w0 << 32) | w1) << o; avail = 64 - (int)o;
  }
  __device__ __forceinline__ uint32_t peek32() const { return (uint32_t)(buf >> 32); }
  __device__ __forceinline__ void consume(int len) {
    buf <<= len; avail -= len;
    if (avail <= 32) { buf |= (unsigned long long)word() << (32 - avail); avail += 32; }
  }
};

// one symbol at bit position p: returns its length (0 = the 32 bits are no code) and the symbol
__device__ __forceinline__ int hfDecodeOne(const HuffFastTables* __restrict__ t, const uint16_t* __restrict__ sLut, uint32_t bits32, int& sym) {
  const uint32_t e = sLut[bits32 >> 20];
  if (e) { sym = (int)(e & 0xff); return (int)(e >> 8); }
  for (int len = max(13, t->minLen); len <= t->maxLen; len++) {
    const uint32_t c = bits32 >> (32 - len);
    const uint32_t d = c - t->first[len];
    if (c >= t->first[len] && d < t->count[len]) { sym = t->syms[t->offset[len] + d]; return len; }
  }
  return 0;
}

// Iteration of the self-synchronising chunk decode.  iter 0: start at the chunk boundary.
__global__ void __launch_bounds__(256) k_huff_chunks(HuffFastArgs a, int iter) {
  __shared__ uint16_t sLut[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) sLut[i] = a.tab->lut[i];
  __syncthreads();
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= a.nChunks) return;
  const unsigned long long* endPrev = (iter & 1) ? a.endA : a.endB;     // written by the previous iteration
  unsigned long long* endCur = (iter & 1) ? a.endB : a.endA;
  const unsigned long long chunkEnd = min((unsigned long long)(t + 1) * HF_CHUNK, a.nBits);
  unsigned long long start;
  if (iter == 0) start = (unsigned long long)t * HF_CHUNK;
  else {
    start = t == 0 ? 0ull : endPrev[t - 1];
    if (start == a.startA[t]) { endCur[t] = endPrev[t]; return; }          // same start as before: same result
  }
  a.startA[t] = start;
  unsigned long long p = start; uint32_t cnt = 0;
  if (p < chunkEnd && p + 192 <= a.nBits) {
    HfReader rd; rd.init(a.stream, p);
    const unsigned long long safeEnd = a.nBits - 192;                      // the reader runs up to two words ahead
    while (p < chunkEnd && p <= safeEnd) {
      int sym; int len = hfDecodeOne(a.tab, sLut, rd.peek32(), sym);
      cnt += len ? 1 : 0;
      len = len ? len : 1;                                                 // a speculative start may see a non-code: skip a bit
      rd.consume(len); p += len;
    }
  }
  while (p < chunkEnd) {                                                   // last words of the stream: position-addressed reads
    if (p + 32 > a.nBits) { p = chunkEnd; break; }                         // inside the read-ahead padding
    int sym; const int len = hfDecodeOne(a.tab, sLut, hfPeek(a.stream, p), sym);
    p += len ? len : 1;
    cnt += len ? 1 : 0;
  }
  endCur[t] = p; a.count[t] = cnt;
  if (iter > 0) *a.changed = 1;
}

// Final pass: symbols into out (IEM_Huffman: raster order, value = symbol - offset) or into the planar delta image.
template <class T>
__global__ void __launch_bounds__(256) k_huff_emit(HuffFastArgs a, const unsigned long long* __restrict__ endFinal, const uint32_t* __restrict__ symBase, uint8_t* __restrict__ out) {
  __shared__ uint16_t sLut[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) sLut[i] = a.tab->lut[i];
  __syncthreads();
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= a.nChunks) return;
  const unsigned long long chunkEnd = min((unsigned long long)(t + 1) * HF_CHUNK, a.nBits);
  unsigned long long p = t == 0 ? 0ull : endFinal[t - 1];
  unsigned long long n = symBase[t];
  bool stop = false;
  // symbols leave four at a time: a word is stored whole once all of its bytes came from this thread, else byte by byte
  const unsigned al = (unsigned)((uintptr_t)out & 3);
  uint32_t wAcc = 0; int have = 0;
  auto put = [&](int sym) {
    const unsigned c = ((unsigned)n + al) & 3u;
    wAcc |= (uint32_t)(sym & 0xff) << (8 * c);
    have++; n++;
    if (c == 3u) {
      if (have == 4) *(uint32_t*)(out + n - 4) = wAcc;
      else for (int q = 0; q < have; q++) out[n - have + q] = (uint8_t)(wAcc >> (8 * (4 - have + q)));
      wAcc = 0; have = 0;
    }
  };
  if (p < chunkEnd && p + 192 <= a.nBits) {
    HfReader rd; rd.init(a.stream, p);
    const unsigned long long safeEnd = a.nBits - 192;
    while (p < chunkEnd && n < a.nSym && p <= safeEnd) {
      int sym; const int len = hfDecodeOne(a.tab, sLut, rd.peek32(), sym);
      if (!len) { atomicOr(a.bad, 1); stop = true; break; }                // no code on the true chain: the serial decoder decides
      put(sym);
      rd.consume(len); p += len;
      if (n == a.nSym) *a.endBit = p;
    }
  }
  while (!stop && p < chunkEnd && n < a.nSym) {
    if (p + 32 > a.nBits) { atomicOr(a.bad, 1); break; }
    int sym; const int len = hfDecodeOne(a.tab, sLut, hfPeek(a.stream, p), sym);
    if (!len) { atomicOr(a.bad, 1); break; }
    put(sym);
    p += len;
    if (n == a.nSym) *a.endBit = p;
  }
  // the bytes of the last, incomplete word: they sit at byte positions c - have + 1 .. c of wAcc, c = class of the last symbol
  for (int q = 0; q < have; q++) {
    const unsigned c = ((unsigned)(n - have + q) + al) & 3u;
    out[n - have + q] = (uint8_t)(wAcc >> (8 * c));
  }
}

// non-delta mode: symbol -> value in place (offset 128 for signed char, Lerc2.cpp:2320)
template <class T>
__global__ void k_huff_values(uint8_t* __restrict__ data, unsigned long long n, int off) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
    data[i] = (uint8_t)(data[i] - off);
}

// delta mode, all pixels valid: running sums.  planes: delta image [D][H][W] of symbols.  col0[d][i] = value of pixel (i, 0).
__global__ void __launch_bounds__(256) k_huff_col0(const uint8_t* __restrict__ planes, int H, int W, int D, int off, uint8_t* __restrict__ col0) {
  // one CTA per depth plane: running sum of the first column down the rows (predictor of (i, 0) is (i-1, 0); of (0, 0): 0)
  __shared__ uint32_t sSum[256];
  const int d = blockIdx.x, tid = threadIdx.x;
  const uint8_t* pl = planes + (size_t)d * H * W;
  const int per = (H + 255) / 256;
  const int i0 = tid * per, i1 = min(H, i0 + per);
  uint32_t s = 0;
  for (int i = i0; i < i1; i++) s += (uint32_t)(uint8_t)(pl[(size_t)i * W] - off);
  sSum[tid] = s;
  __syncthreads();
  if (tid == 0) { uint32_t run = 0; for (int k = 0; k < 256; k++) { const uint32_t v = sSum[k]; sSum[k] = run; run += v; } }
  __syncthreads();
  uint32_t run = sSum[tid];
  for (int i = i0; i < i1; i++) { run += (uint32_t)(uint8_t)(pl[(size_t)i * W] - off); col0[(size_t)d * H + i] = (uint8_t)run; }
}

// one CTA per image row: per depth plane a running sum along the row starting from col0, written pixel-interleaved
template <class T>
__global__ void __launch_bounds__(256) k_huff_rows(const uint8_t* __restrict__ planes, const uint8_t* __restrict__ col0, int H, int W, int D, int off, T* __restrict__ data) {
  constexpr int SEG = 4096;                                   // pixels per pass
  __shared__ uint32_t sSum[256];
  __shared__ uint32_t sCarry;
  const int i = blockIdx.x, tid = threadIdx.x;
  for (int d = 0; d < D; d++) {
    const uint8_t* row = planes + ((size_t)d * H + i) * W;
    if (tid == 0) sCarry = col0[(size_t)d * H + i];           // value of pixel (i, 0)
    __syncthreads();
    for (int j0 = 0; j0 < W; j0 += SEG) {
      const int per = SEG / 256;
      const int a0 = j0 + tid * per, a1 = min(W, a0 + per);
      uint32_t s = 0;
      for (int j = a0; j < a1; j++) if (j > 0) s += (uint32_t)(uint8_t)(row[j] - off);
      sSum[tid] = s;
      __syncthreads();
      if (tid == 0) { uint32_t run = sCarry; for (int k = 0; k < 256; k++) { const uint32_t v = sSum[k]; sSum[k] = run; run += v; } sCarry = run; }
      __syncthreads();
      uint32_t run = sSum[tid];
      for (int j = a0; j < a1; j++) {
        if (j > 0) run += (uint32_t)(uint8_t)(row[j] - off);
        data[((size_t)i * W + j) * D + d] = (T)(uint8_t)run;
      }
      __syncthreads();
    }
  }
}

// The same for D = 1 or 3 planes, rows of a multiple of 16 pixels and 16-byte aligned buffers (BASELINE config 4: RGB): every thread
// takes 16 pixels of ALL planes (one 16-byte load per plane), sums them as packed bytes (the running sum is modulo 256 anyway),
// the totals go through a warp scan and one shared-memory hop per 4096-pixel pass, and the pixel-interleaved result leaves as
// 16-byte stores.
template <int DD>
__global__ void __launch_bounds__(256) k_huff_rows_vec(const uint8_t* __restrict__ planes, const uint8_t* __restrict__ col0, int H, int W, int off, uint8_t* __restrict__ data) {
  __shared__ uint32_t sWarp[8][DD];
  __shared__ uint32_t sCarry[DD];
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t off4 = (uint32_t)off * 0x01010101u;
  if (tid < DD) sCarry[tid] = 0;
  __syncthreads();
  for (int j0 = 0; j0 < W; j0 += 4096) {
    const int a0 = j0 + tid * 16;
    const bool act = a0 < W;
    uint32_t x[DD][4], tot[DD];
#pragma unroll
    for (int d = 0; d < DD; d++) {
      uint4 v = make_uint4(off4, off4, off4, off4);
      if (act) v = __ldg((const uint4*)(planes + ((size_t)d * H + i) * W + a0));
      x[d][0] = __vsub4(v.x, off4); x[d][1] = __vsub4(v.y, off4); x[d][2] = __vsub4(v.z, off4); x[d][3] = __vsub4(v.w, off4);
      if (a0 == 0) x[d][0] = (x[d][0] & 0xffffff00u) | (uint32_t)col0[(size_t)d * H + i];      // pixel (i, 0): its value, not a delta (predictor = pixel above)
      uint32_t run = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {                                       // running sums of the 16 bytes
        uint32_t w = x[d][k];
        w = __vadd4(w, w << 8); w = __vadd4(w, w << 16);
        w = __vadd4(w, run * 0x01010101u);
        run = w >> 24;
        x[d][k] = w;
      }
      tot[d] = run;
    }
    // exclusive prefix of the threads' totals (mod 256) inside the warp, then across the warps and the passes
#pragma unroll
    for (int d = 0; d < DD; d++) {
      uint32_t inc = tot[d];
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) { const uint32_t o = __shfl_up_sync(FULL, inc, m); if (lane >= m) inc += o; }
      if (lane == 31) sWarp[warp][d] = inc;
      tot[d] = inc - tot[d];
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < DD; d++) {
      uint32_t base = sCarry[d];
      for (int w = 0; w < warp; w++) base += sWarp[w][d];
      const uint32_t add = ((base + tot[d]) & 0xffu) * 0x01010101u;
#pragma unroll
      for (int k = 0; k < 4; k++) x[d][k] = __vadd4(x[d][k], add);
    }
    if (act) {
      uint8_t* dst = data + ((size_t)i * W + a0) * DD;
      if (DD == 1) *(uint4*)dst = make_uint4(x[0][0], x[0][1], x[0][2], x[0][3]);
      else {
        uint32_t o[12];
#pragma unroll
        for (int k = 0; k < 4; k++) {                                     // 4 pixels of 3 planes -> 3 interleaved words
          const uint32_t A = x[0][k], B = x[1 % DD][k], C = x[2 % DD][k];
          o[3 * k + 0] = __byte_perm(__byte_perm(A, B, 0x1040), C, 0x3410);      // A0 B0 C0 A1
          o[3 * k + 1] = __byte_perm(__byte_perm(B, C, 0x2051), A, 0x3610);      // B1 C1 A2 B2
          o[3 * k + 2] = __byte_perm(__byte_perm(C, A, 0x3072), B, 0x3710);      // C2 A3 B3 C3
        }
        uint4* d4 = (uint4*)dst;
        d4[0] = make_uint4(o[0], o[1], o[2], o[3]); d4[1] = make_uint4(o[4], o[5], o[6], o[7]); d4[2] = make_uint4(o[8], o[9], o[10], o[11]);
      }
    }
    __syncthreads();
    if (tid == 255) {
#pragma unroll
      for (int d = 0; d < DD; d++) { uint32_t base = sCarry[d]; for (int w = 0; w < 8; w++) base += sWarp[w][d]; sCarry[d] = base & 0xffu; }
    }
    __syncthreads();
  }
}

}  // namespace lerc
