// lerc_fpl_decode.cuh -- decoder of the reference's lossless float / double codec "FPL" (image mode IEM_DeltaDeltaHuffman,
// Lerc2 v6, maxZError == 0).  Included by lerc_decode.cu.
//
// Blob layout behind the image-mode byte (fpl_Lerc2Ext.cpp:738-866, DecodeHuffmanFltSlice): one predictor code byte, then for
// each of the 4 (float) / 8 (double) BYTE PLANES of the values {plane index, byte-delta level, compressed size u32, compressed
// bytes}.  A plane is coded as one of (fpl_EsriHuffman.cpp:243, :453-558): 0 = canonical Huffman (same code table and MSB-first
// bit stream as the 8-bit Huffman path, incl. the read-ahead word), 1 = one repeated value, 2 = stored, 3 = PackBits runs.
// Decoding = plane bytes -> `level` running sums mod 256 (restoreSequence, fpl_Lerc2Ext.cpp:133-169) -> planes interleaved into
// 32 / 64-bit units -> predictor undone: running "sums" along rows, for predictor 2 first down the columns, with the codec's
// split addition (mantissa field and exponent/sign field wrap separately: ADD32_BIT_FLT / ADD64_BIT_DBL, fpl_UnitTypes.cpp:98-156;
// both fields are plain modular adds, so the operator is associative and the sums are parallel scans) -> floats: exponent and
// sign moved back (undo_moveBits2Front, fpl_UnitTypes.cpp:51-63).  The WHOLE array is coded, invalid pixels included.
// nDepth > 1 is one slice of nDepth columns and nCols * nRows rows (fpl_Lerc2Ext.cpp:725-736).
//
// Kernels: the Huffman planes reuse the parallel self-synchronising decoder of lerc_huffman_fast.cuh (serial kernel as the
// fallback); byte running sums are cub::DeviceScan::InclusiveSum over uint8; k_fpl_scan is a one-CTA-per-sequence scan with the
// split addition (rows: coalesced; columns: strided -- first version, not tuned); k_fpl_packbits walks the run headers with one
// lane and fills with the warp.
#pragma once
#include <functional>

namespace lerc {

enum { FPLF_BAD = 1 };

__device__ __forceinline__ uint32_t fplAdd(uint32_t a, uint32_t b) {
  return ((a + b) & 0x007FFFFFu) | (((((a >> 23) & 0x1FFu) + ((b >> 23) & 0x1FFu)) & 0x1FFu) << 23);
}
__device__ __forceinline__ unsigned long long fplAdd(unsigned long long a, unsigned long long b) {
  return ((a + b) & 0x000FFFFFFFFFFFFFull) | (((((a >> 52) & 0xFFFull) + ((b >> 52) & 0xFFFull)) & 0xFFFull) << 52);
}

// PackBits (fpl_EsriHuffman.cpp:37-75): header b <= 127: b + 1 literal bytes follow; else one byte repeated b - 126 times.
__global__ void k_fpl_packbits(const uint8_t* __restrict__ src, unsigned long long size, uint8_t* __restrict__ out, unsigned long long n, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  unsigned long long i = 0, cur = 0;
  bool bad = false;
  while (i < size) {
    const int b = src[i];
    if (b <= 127) {
      const unsigned long long c = (unsigned long long)b + 1;
      if (cur + (unsigned long long)b >= n || i + 1 + c > size) { bad = true; break; }
      for (unsigned long long k = lane; k < c; k += 32) out[cur + k] = src[i + 1 + k];
      cur += c; i += 1 + c;
    } else {
      const unsigned long long c = (unsigned long long)b - 127 + 1;
      if (cur + (unsigned long long)b - 127 >= n || i + 1 >= size) { bad = true; break; }
      const uint8_t v = src[i + 1];
      for (unsigned long long k = lane; k < c; k += 32) out[cur + k] = v;
      cur += c; i += 2;
    }
  }
  if (lane == 0 && (bad || cur != n)) atomicOr(status, FPLF_BAD);
}

// unit i, byte idx[b] <- plane b, element i
template <class U>
__global__ void k_fpl_assemble(const uint8_t* __restrict__ planes, unsigned long long n, int idx0, int idx1, int idx2, int idx3, int idx4, int idx5, int idx6, int idx7,
                               U* __restrict__ out) {
  const int idx[8] = {idx0, idx1, idx2, idx3, idx4, idx5, idx6, idx7};
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    U v = 0;
#pragma unroll
    for (int b = 0; b < (int)sizeof(U); b++) v |= (U)planes[(unsigned long long)b * n + i] << (8 * idx[b]);
    out[i] = v;
  }
}

// y[k] = y[k] (+) y[k - 1] for k = 1 .. len - 1 along every one of `outer` sequences; element k of sequence o lives at
// data[o * outerStride + k * stride].  One CTA per sequence, 256 threads x 8 elements per pass, running carry between passes.
template <class U>
__global__ void __launch_bounds__(256) k_fpl_scan(U* __restrict__ data, unsigned long long outer, unsigned long long outerStride, unsigned long long len, unsigned long long stride) {
  constexpr int PER = 8;
  __shared__ U sWarp[8];
  __shared__ U sCarry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (unsigned long long o = blockIdx.x; o < outer; o += gridDim.x) {
    U* seq = data + o * outerStride;
    if (tid == 0) sCarry = 0;
    __syncthreads();
    for (unsigned long long base = 0; base < len; base += 256ull * PER) {
      U v[PER];
      U run = 0;
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const unsigned long long e = base + (unsigned long long)tid * PER + k;
        v[k] = e < len ? seq[e * stride] : (U)0;
        run = fplAdd(run, v[k]); v[k] = run;                      // inclusive scan of this thread's elements
      }
      U inc = run;                                                // inclusive scan of the thread totals inside the warp
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) { const U up = __shfl_up_sync(FULL, inc, m); if (lane >= m) inc = fplAdd(inc, up); }
      if (lane == 31) sWarp[warp] = inc;
      __syncthreads();
      U before = sCarry;                                          // everything in front of this thread: carry, earlier warps, earlier lanes
      for (int w2 = 0; w2 < warp; w2++) before = fplAdd(before, sWarp[w2]);
      const U excl = __shfl_up_sync(FULL, inc, 1);
      if (lane > 0) before = fplAdd(before, excl);
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const unsigned long long e = base + (unsigned long long)tid * PER + k;
        if (e < len) seq[e * stride] = fplAdd(before, v[k]);
      }
      __syncthreads();
      if (tid == 255) sCarry = fplAdd(before, run);
      __syncthreads();
    }
  }
}

// undo_moveBits2Front (fpl_UnitTypes.cpp:51-63): [exponent 8 | sign 1 | mantissa 23] -> IEEE [sign | exponent | mantissa]
__global__ void k_fpl_unmove(uint32_t* __restrict__ data, unsigned long long n) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t a = data[i];
    data[i] = (a & 0x007FFFFFu) | (((a >> 24) & 0xFFu) << 23) | (((a >> 23) & 1u) << 31);
  }
}

namespace {

// fetch(offset, length, dst): host view of this band's blob.  dBlob: the same bytes on the device.  pos: first byte behind the
// image-mode byte.  dOut: the band's pixels (all of them are written).
template <class T>
ErrCode decodeFpl(Context* ctx, const std::function<bool(size_t, size_t, void*)>& fetch, const uint8_t* dBlob, size_t pos, size_t blobSize,
                  int nCols, int nRows, int nDepth, void* dOut, int* dStatus) {
  using U = typename std::conditional<sizeof(T) == 8, unsigned long long, uint32_t>::type;
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "FPL codes float and double only");
  cudaStream_t st = ctx->stream;
  const unsigned long long cols = nDepth == 1 ? (unsigned long long)nCols : (unsigned long long)nDepth;
  const unsigned long long rows = nDepth == 1 ? (unsigned long long)nRows : (unsigned long long)nCols * (unsigned long long)nRows;
  const unsigned long long n = cols * rows;
  constexpr int unit = (int)sizeof(T);
  if (n == 0 || n > 0x7fffffffull || pos + 1 > blobSize) return Failed;
  uint8_t pred = 0;
  if (!fetch(pos, 1, &pred) || pred > 2) return Failed;
  pos += 1;
  uint8_t* dPlanes = (uint8_t*)ctx->arena.alloc((size_t)n * unit + 16);
  if (!dPlanes) return Failed;
  int idx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int fillGrid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((n + 255) / 256, 148ull * 32));
  for (int b = 0; b < unit; b++) {
    uint8_t hdr[6];
    if (pos + 6 > blobSize || !fetch(pos, 6, hdr)) return Failed;
    const int planeIdx = hdr[0], level = hdr[1];
    uint32_t csize; std::memcpy(&csize, hdr + 2, 4);
    pos += 6;
    if (planeIdx >= unit || level > 5 || csize < 1 || pos + (size_t)csize > blobSize) return Failed;
    idx[b] = planeIdx;
    uint8_t* dPlane = dPlanes + (size_t)b * n;
    uint8_t head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!fetch(pos, std::min<size_t>(6, csize), head)) return Failed;
    switch (head[0]) {
      case 1: {                                                     // one repeated value
        uint32_t cnt; std::memcpy(&cnt, head + 2, 4);
        if (csize < 6 || (unsigned long long)cnt != n) return Failed;
        cudaMemsetAsync(dPlane, head[1], (size_t)n, st);
        break;
      }
      case 2:                                                       // stored
        if ((unsigned long long)csize < 1 + n) return Failed;
        cudaMemcpyAsync(dPlane, dBlob + pos + 1, (size_t)n, cudaMemcpyDeviceToDevice, st);
        break;
      case 3:                                                       // PackBits
        LERC_LAUNCH(ctx, k_fpl_packbits, 1, 32, 0, dBlob + pos + 1, (unsigned long long)csize - 1, dPlane, n, dStatus);
        break;
      case 0: {                                                     // Huffman
        std::vector<uint8_t> tb(std::min<size_t>(2048, (size_t)csize - 1));
        if (tb.empty() || !fetch(pos + 1, tb.size(), tb.data())) return Failed;
        HuffmanTable t;
        const size_t used = t.read(tb.data(), tb.size(), 5);
        if (!used || 1 + used > (size_t)csize) return Failed;
        const uint8_t* dStream = dBlob + pos + 1 + used;
        const size_t streamLen = (size_t)csize - 1 - used;
        const int rc = decodeHuffmanFast<uint8_t>(ctx, t, dStream, streamLen, (int)rows, (int)cols, 1, false, dPlane);
        if (rc < 0) return Failed;
        if (rc == 0) {
          HuffDecArgs ha;
          ha.stream = dStream; ha.streamLen = (unsigned long long)streamLen; ha.bits = nullptr; ha.H = (int)rows; ha.W = (int)cols; ha.D = 1;
          ha.delta = 0; ha.allValidImage = 1;
          std::memcpy(ha.len, t.len, sizeof ha.len); std::memcpy(ha.code, t.code, sizeof ha.code);
          ha.data = dPlane; ha.status = dStatus;
          LERC_LAUNCH(ctx, k_huffman_decode_seq<uint8_t>, 1, 32, 0, ha);
        }
        break;
      }
      default: return Failed;
    }
    pos += csize;
    for (int l = level; l > 0; l--) {                               // restoreSequence: running sums mod 256 over elements l - 1 .. n - 1
      if ((unsigned long long)l > n) continue;
      LaunchScope scope(ctx, "cub::InclusiveSum<u8>");
      size_t tmpBytes = 0;
      cub::DeviceScan::InclusiveSum(nullptr, tmpBytes, dPlane + (l - 1), dPlane + (l - 1), (int)(n - (unsigned long long)(l - 1)), st);
      void* tmp = ctx->arena.alloc(tmpBytes ? tmpBytes : 16);
      if (!tmp) return Failed;
      cub::DeviceScan::InclusiveSum(tmp, tmpBytes, dPlane + (l - 1), dPlane + (l - 1), (int)(n - (unsigned long long)(l - 1)), st);
      ctx->kernelLaunches += 2;
    }
  }
  U* out = (U*)dOut;
  LERC_LAUNCH(ctx, k_fpl_assemble<U>, fillGrid, 256, 0, (const uint8_t*)dPlanes, n, idx[0], idx[1], idx[2], idx[3], idx[4], idx[5], idx[6], idx[7], out);
  if (pred == 2 && rows > 1) LERC_LAUNCH(ctx, k_fpl_scan<U>, (unsigned)std::min<unsigned long long>(cols, 148ull * 8), 256, 0, out, cols, 1ull, rows, cols);     // down the columns
  if (pred >= 1 && cols > 1) LERC_LAUNCH(ctx, k_fpl_scan<U>, (unsigned)std::min<unsigned long long>(rows, 148ull * 8), 256, 0, out, rows, cols, cols, 1ull);     // along the rows
  if (sizeof(T) == 4) LERC_LAUNCH(ctx, k_fpl_unmove, fillGrid, 256, 0, (uint32_t*)out, n);
  return cudaOk(cudaGetLastError(), "decodeFpl") ? Ok : Failed;
}

}  // namespace
}  // namespace lerc
