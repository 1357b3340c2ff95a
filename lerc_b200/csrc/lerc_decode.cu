// lerc_decode.cu -- Lerc2 (v3..v6) band decoder on the GPU.
//
// Pipeline of one band (reference: Lerc2::Decode Lerc2.cpp:577-694):
//   Fletcher-32 verify  ->  mask (RLE decode | all / none / previous band)  ->  zero fill  ->
//   const image | per-depth const | one-sweep raw | 8-bit Huffman | micro-block stream:
//        block boundary discovery (the stream has no index)  ->  per-block unpack /
//        dequantise / clamp / cast.
// Reference citations relative to /root/reference/src/LercLib.
#include "lerc_device.cuh"
#include "lerc_kernels.h"
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cmath>

namespace lerc {

enum { DECF_BAD_STREAM = 1, DECF_BAD_CHECKSUM = 2, DECF_BAD_MASK = 4 };

struct DecTileArgs {
  const uint8_t* stream; unsigned long long streamLen;   // micro-block stream (after the flag bytes) and its length
  const uint8_t* bits;                                    // never null on decode (all-valid masks are 0xff)
  int nRows, nCols, nDepth, mb, nTx, nTy, dt, version, allValidImage;
  double maxZErr, zMaxHdr;
  const double* zMaxVec;                                  // per depth (version >= 4 && nDepth > 1) or nullptr
  uint32_t* blockOff;                                     // [nBlocks] stream offset of each block's first unit
  void* data;
  int* status;
};

// number of valid pixels in block (ty, tx)
__device__ inline int blockValidCount(const DecTileArgs& a, int i0, int j0, int h, int w) {
  if (a.allValidImage) return h * w;
  int n = 0;
  for (int r = 0; r < h; r++) {
    const long long k0 = (long long)(i0 + r) * a.nCols + j0;
    for (int c = 0; c < w; c++) n += maskBit(a.bits, k0 + c) ? 1 : 0;
  }
  return n;
}

// Length in bytes of the unit (one block, one depth) whose first byte is p[0]; p must expose 16 readable bytes.
// Follows ReadTile (Lerc2.cpp:2025-2230) and BitStuffer2::Decode (BitStuffer2.cpp:159-258).  0 = malformed.
__device__ inline unsigned unitLength(const uint8_t* p, int dt, int version, int rawCount, int maxCount) {
  const unsigned flag = p[0], mode = flag & 3;
  const bool diff = version >= 5 && (flag & 4);
  if (mode == 2) return 1;
  if (mode == 0) return diff ? 0 : 1 + (unsigned)rawCount * (unsigned)dtSize(dt);
  const int dtUsed = offsetTypeFromCode((diff && dt < DT_Float) ? DT_Int : dt, (int)(flag >> 6));
  if (dtUsed == DT_Undefined) return 0;
  const int osz = dtSize(dtUsed);
  if (mode == 3) return 1 + osz;
  const unsigned b = p[1 + osz], nb = b & 31, lut = (b >> 5) & 1, code = b >> 6;
  const int cb = code == 0 ? 4 : 3 - (int)code;
  if (cb <= 0) return 0;
  const unsigned n = (unsigned)loadBytesLE(p + 2 + osz, cb);
  if (n > (unsigned)maxCount) return 0;
  unsigned len = 2 + osz + cb;
  if (!lut) {
    if (nb > 0) { if (n == 0) return 0; len += packedBytes(n, nb); }
  } else {
    if (nb == 0 || n == 0) return 0;
    const int nLut = (int)p[len] - 1;
    if (nLut < 1) return 0;
    len += 1 + packedBytes(nLut, nb) + packedBytes(n, bitLength((uint32_t)nLut));
  }
  return len;
}

// ------------------------------------------------------------------------------------------------
// block boundary discovery, sequential form: exact for every stream (masks, edge blocks, raw blocks).
// One CTA stages the stream through shared memory 8 KB at a time; thread 0 hops from header to header.
__global__ void k_walk_units(DecTileArgs a) {
  constexpr int CH = 8192;
  __shared__ uint8_t buf[CH + 32];
  __shared__ unsigned long long sCur;
  __shared__ int sBlk, sDepth, sBad;
  const int nBlocks = a.nTx * a.nTy;
  if (threadIdx.x == 0) { sCur = 0; sBlk = 0; sDepth = 0; sBad = 0; }
  __syncthreads();
  const int pattern = a.version >= 5 ? 14 : 15;
  while (sBlk < nBlocks && !sBad) {
    const unsigned long long base = sCur;
    for (int i = threadIdx.x; i < CH + 32; i += blockDim.x) buf[i] = (base + i < a.streamLen) ? a.stream[base + i] : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long cur = base;
      int blk = sBlk, d = sDepth;
      while (blk < nBlocks && cur + 16 <= base + CH + 32) {
        if (cur >= a.streamLen) { sBad = 1; break; }
        const int ty = blk / a.nTx, tx = blk - ty * a.nTx, i0 = ty * a.mb, j0 = tx * a.mb;
        const int h = (i0 + a.mb > a.nRows) ? a.nRows - i0 : a.mb, w = (j0 + a.mb > a.nCols) ? a.nCols - j0 : a.mb;
        const uint8_t* p = buf + (cur - base);
        if ((((p[0] >> 2) & pattern) != ((j0 >> 3) & pattern)) || (a.version >= 5 && (p[0] & 4) && d == 0)) { sBad = 1; break; }
        const int rawCount = ((p[0] & 3) == 0) ? blockValidCount(a, i0, j0, h, w) : 0;
        const unsigned len = unitLength(p, a.dt, a.version, rawCount, h * w);
        if (!len || cur + len > a.streamLen) { sBad = 1; break; }
        if (d == 0) a.blockOff[blk] = (uint32_t)cur;
        cur += len;
        if (++d == a.nDepth) { d = 0; blk++; }
      }
      sCur = cur; sBlk = blk; sDepth = d;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && sBad) atomicOr(a.status, DECF_BAD_STREAM);
}

// The walk's checks, block-parallel, for offsets found speculatively (k_decode_stream<T, true>): every block's units, parsed with the
// block's true size and valid count, must carry the right integrity bits and end exactly where the next block starts.  status[2] != 0:
// the offsets are not the serial parse (the serial walk then decides).
__global__ void k_verify_offsets(DecTileArgs a) {
  const int nBlocks = a.nTx * a.nTy;
  const int pattern = a.version >= 5 ? 14 : 15;
  for (int blk = blockIdx.x * blockDim.x + threadIdx.x; blk < nBlocks; blk += gridDim.x * blockDim.x) {
    const int ty = blk / a.nTx, tx = blk - ty * a.nTx, i0 = ty * a.mb, j0 = tx * a.mb;
    const int h = (i0 + a.mb > a.nRows) ? a.nRows - i0 : a.mb, w = (j0 + a.mb > a.nCols) ? a.nCols - j0 : a.mb;
    unsigned long long cur = a.blockOff[blk];
    bool bad = blk == 0 && cur != 0;
    for (int d = 0; d < a.nDepth && !bad; d++) {
      if (cur >= a.streamLen) { bad = true; break; }
      uint8_t hb[16];
      for (int k = 0; k < 16; k++) hb[k] = cur + k < a.streamLen ? a.stream[cur + k] : 0;
      if ((((hb[0] >> 2) & pattern) != ((j0 >> 3) & pattern)) || (a.version >= 5 && (hb[0] & 4) && d == 0)) { bad = true; break; }
      const int rawCount = ((hb[0] & 3) == 0) ? blockValidCount(a, i0, j0, h, w) : 0;
      const unsigned len = unitLength(hb, a.dt, a.version, rawCount, h * w);
      if (!len || cur + len > a.streamLen) { bad = true; break; }
      cur += len;
    }
    if (!bad && blk + 1 < nBlocks && cur != (unsigned long long)a.blockOff[blk + 1]) bad = true;
#ifdef LERC_CUSIM
    if (bad && std::getenv("DS_DEBUG3")) {
      std::fprintf(stderr, "      verify: block %d (ty %d tx %d) off %u next %u parsed end %llu valid %d bytes", blk, ty, tx, a.blockOff[blk], blk + 1 < nBlocks ? a.blockOff[blk + 1] : 0u, cur, blockValidCount(a, i0, j0, h, w));
      for (int k = 0; k < 20; k++) std::fprintf(stderr, " %02x", a.stream[a.blockOff[blk] + k]);
      std::fprintf(stderr, "\n");
    }
#endif
    if (bad) atomicOr(a.status + 2, 1);
  }
}

// ------------------------------------------------------------------------------------------------
// micro-block decoder: one warp per block position, all depths               Lerc2.cpp:2025-2230
template <class T> __device__ __forceinline__ T castClamped(double z, double zMax) {
  const double v = z < zMax ? z : zMax;             // std::min(z, zMax), Lerc2.cpp:2160
  return (T)v;
}

// codec version 2 (MSB-first words, BitStuffer2.cpp:352-425): element e of n
__device__ inline uint32_t extractBitsV2(const uint8_t* __restrict__ payload, uint32_t payloadLen, uint32_t e, int nb, uint32_t n) {
  uint32_t x = 0;
  for (int t = 0; t < nb; t++) {
    const uint32_t sb = e * (uint32_t)nb + (uint32_t)t;
    const int k = v2ByteOfStreamBit(sb, n, nb);
    const uint32_t bit = (k >= 0 && (uint32_t)k < payloadLen) ? ((payload[k] >> (7 - (sb & 7))) & 1u) : 0u;
    x = (x << 1) | bit;
  }
  return x;
}

__device__ inline uint32_t extractBits(const uint8_t* __restrict__ payload, uint32_t payloadLen, uint32_t e, int nb) {
  const unsigned long long bit = (unsigned long long)e * nb;
  const uint32_t k0 = (uint32_t)(bit >> 3); const int sh = (int)(bit & 7);
  unsigned long long x = 0;
  const int need = (sh + nb + 7) >> 3;
  for (int k = 0; k < need && k0 + k < payloadLen; k++) x |= (unsigned long long)payload[k0 + k] << (8 * k);
  return (uint32_t)(x >> sh) & (nb == 32 ? 0xffffffffu : ((1u << nb) - 1));
}

template <class T>
__global__ void k_tiles_decode(DecTileArgs a) {
  __shared__ uint32_t sLut[8][256];
  const int warpsPerCta = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* data = (T*)a.data;
  const int nBlocks = a.nTx * a.nTy;
  const double invScale = __dmul_rn(2.0, a.maxZErr);
  for (int blk = blockIdx.x * warpsPerCta + warp; blk < nBlocks; blk += gridDim.x * warpsPerCta) {
    const int ty = blk / a.nTx, tx = blk - ty * a.nTx, i0 = ty * a.mb, j0 = tx * a.mb;
    const int h = (i0 + a.mb > a.nRows) ? a.nRows - i0 : a.mb, w = (j0 + a.mb > a.nCols) ? a.nCols - j0 : a.mb;
    const int cells = h * w;
    const uint8_t* p = a.stream + a.blockOff[blk];
    bool bad = false;
    for (int d = 0; d < a.nDepth && !bad; d++) {
      const unsigned flag = p[0], mode = flag & 3;
      const bool diff = a.version >= 5 && (flag & 4);
      const double zMax = a.zMaxVec ? a.zMaxVec[d] : a.zMaxHdr;
      double offset = 0;
      int nb = 0, cb = 0, nLut = 0, nbIdx = 0; unsigned n = 0; bool lut = false;
      const uint8_t* payload = p + 1; uint32_t payloadLen = 0;
      unsigned unitLen = 1;
      if (mode == 0) {
        // raw values of the valid pixels follow the flag byte
      } else if (mode != 2) {
        const int dtUsed = offsetTypeFromCode((diff && a.dt < DT_Float) ? DT_Int : a.dt, (int)(flag >> 6));
        const int osz = dtSize(dtUsed);
        offset = offsetFromBits(loadBytesLE(p + 1, osz), dtUsed);
        unitLen = 1 + osz;
        if (mode == 1) {
          const unsigned b = p[1 + osz];
          nb = b & 31; lut = (b >> 5) & 1; const unsigned code = b >> 6; cb = code == 0 ? 4 : 3 - (int)code;
          n = (unsigned)loadBytesLE(p + 2 + osz, cb);
          unitLen = 2 + osz + cb;
          if (lut) {
            nLut = (int)p[unitLen] - 1; unitLen += 1;
            const uint8_t* lp = p + unitLen; const uint32_t lutLen = packedBytes(nLut, nb);
            if (lane == 0) sLut[warp][0] = 0;
            for (int i = lane; i < nLut; i += 32) sLut[warp][1 + i] = a.version >= 3 ? extractBits(lp, lutLen, i, nb) : extractBitsV2(lp, lutLen, i, nb, (uint32_t)nLut);
            __syncwarp();
            unitLen += lutLen;
            nbIdx = bitLength((uint32_t)nLut);
            payload = p + unitLen; payloadLen = packedBytes(n, nbIdx); unitLen += payloadLen;
          } else {
            payload = p + unitLen; payloadLen = nb ? packedBytes(n, nb) : 0; unitLen += payloadLen;
          }
        }
      }
      // A full count addresses every cell of the block, mask or not (Lerc2.cpp:2148); otherwise values map to the valid pixels in order.
      const bool allCells = (mode == 1) && (n == (unsigned)cells);
      int rank = 0;
      for (int base = 0; base < cells; base += 32) {
        const int c = base + lane;
        bool valid = false; long long m = 0;
        if (c < cells) {
          const int r = c / w, col = c - r * w;
          const long long k = (long long)(i0 + r) * a.nCols + (j0 + col);
          valid = allCells ? true : maskBit(a.bits, k);
          m = k * a.nDepth + d;
        }
        const unsigned bal = __ballot_sync(FULL, valid);
        const int e = rank + __popc(bal & ((1u << lane) - 1));
        rank += __popc(bal);
        if (!valid) continue;
        if (mode == 2) data[m] = diff ? data[m - 1] : (T)0;
        else if (mode == 0) {
          T v; uint8_t* vb = (uint8_t*)&v; const uint8_t* src = p + 1 + (size_t)e * sizeof(T);
          for (int b = 0; b < (int)sizeof(T); b++) vb[b] = src[b];
          data[m] = v;
        } else if (mode == 3) {
          if (!diff) data[m] = (T)offset;
          else data[m] = castClamped<T>(__dadd_rn(offset, (double)data[m - 1]), zMax);
        } else {
          if ((unsigned)e >= n) { bad = true; continue; }
          uint32_t q = 0;
          if (lut) { const uint32_t idx = a.version >= 3 ? extractBits(payload, payloadLen, e, nbIdx) : extractBitsV2(payload, payloadLen, e, nbIdx, n); if (idx > (uint32_t)nLut) { bad = true; continue; } q = sLut[warp][idx]; }
          else if (nb) q = a.version >= 3 ? extractBits(payload, payloadLen, e, nb) : extractBitsV2(payload, payloadLen, e, nb, n);
          double z = __dadd_rn(offset, __dmul_rn((double)q, invScale));
          if (diff) z = __dadd_rn(z, (double)data[m - 1]);
          data[m] = castClamped<T>(z, zMax);
        }
      }
      if (mode == 0) unitLen = 1 + (unsigned)rank * (unsigned)sizeof(T);
      bad = __any_sync(FULL, bad);
      p += unitLen;
      __syncwarp();
    }
    if (bad && lane == 0) atomicOr(a.status, DECF_BAD_STREAM);
  }
}

// ------------------------------------------------------------------------------------------------
// whole-image fills                                                           Lerc2.cpp:2681-2721, :1368-1400
template <class T>
__global__ void k_fill_const(T* __restrict__ data, const uint8_t* __restrict__ bits, long long nPix, int nDepth, double z0, const double* __restrict__ perDepth) {
  const long long nElem = nPix * nDepth;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nElem; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / nDepth; const int d = (int)(e - k * nDepth);
    if (maskBit(bits, k)) data[e] = (T)(perDepth ? perDepth[d] : z0);
  }
}

template <class T>
__global__ void k_one_sweep_scatter(T* __restrict__ data, const uint8_t* __restrict__ bits, const uint32_t* __restrict__ chunkBase,
                                    long long nPix, int nDepth, const uint8_t* __restrict__ src) {
  const int lane = threadIdx.x & 31, warpsPerCta = blockDim.x >> 5;
  const int nChunks = (int)((nPix + 1023) >> 10);
  const size_t len = (size_t)nDepth * sizeof(T);
  for (int c = blockIdx.x * warpsPerCta + (threadIdx.x >> 5); c < nChunks; c += gridDim.x * warpsPerCta) {
    unsigned long long rank = chunkBase[c];
    for (int step = 0; step < 32; step++) {
      const long long k = (long long)c * 1024 + step * 32 + lane;
      const bool valid = k < nPix && maskBit(bits, k);
      const unsigned m = __ballot_sync(FULL, valid);
      if (valid) {
        const uint8_t* s = src + (rank + __popc(m & ((1u << lane) - 1))) * len;
        uint8_t* dst = (uint8_t*)(data + k * nDepth);
        for (size_t b = 0; b < len; b++) dst[b] = s[b];
      }
      rank += __popc(m);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 8-bit Huffman decode, sequential form (one thread walks the single bit stream)   Lerc2.cpp:2472-2606
// The code table is explicit (length, code) pairs; they are loaded into a binary trie in shared memory.
struct HuffDecArgs {
  const uint8_t* stream; unsigned long long streamLen;     // bit stream (after the code table)
  const uint8_t* bits; int H, W, D, delta, allValidImage;
  uint16_t len[256]; uint32_t code[256];
  void* data; int* status;
};

template <class T>
__global__ void k_huffman_decode_seq(HuffDecArgs a) {
  __shared__ short kid[1024][2];
  __shared__ short sym[1024];
  __shared__ int ok;
  if (threadIdx.x != 0) return;
  int used = 1; kid[0][0] = kid[0][1] = -1; sym[0] = -1; ok = 1;
  for (int s = 0; s < 256 && ok; s++) {
    if (!a.len[s]) continue;
    int cur = 0;
    for (int b = a.len[s] - 1; b >= 0; b--) {
      const int bit = (a.code[s] >> b) & 1;
      if (sym[cur] >= 0) { ok = 0; break; }
      if (kid[cur][bit] < 0) { if (used >= 1024) { ok = 0; break; } kid[used][0] = kid[used][1] = -1; sym[used] = -1; kid[cur][bit] = (short)used++; }
      cur = kid[cur][bit];
    }
    if (ok) { if (kid[cur][0] >= 0 || kid[cur][1] >= 0 || sym[cur] >= 0) ok = 0; else sym[cur] = (short)s; }
  }
  T* data = (T*)a.data;
  const int off = PixelTraits<T>::code == DT_Char ? 128 : 0;
  unsigned long long pos = 0; const unsigned long long limit = (a.streamLen & ~3ull) * 8;
  uint32_t word = 0;
  auto nextSymbol = [&]() -> int {
    int cur = 0;
    while (sym[cur] < 0) {
      if (pos >= limit) { ok = 0; return 0; }
      if ((pos & 31) == 0 || pos == 0) word = (uint32_t)loadBytesLE(a.stream + (pos >> 5) * 4, 4);
      const int bit = (word >> (31 - (pos & 31))) & 1;
      pos++;
      const int nxt = kid[cur][bit];
      if (nxt < 0) { ok = 0; return 0; }
      cur = nxt;
    }
    return sym[cur];
  };
  // (the cached word is refreshed whenever pos crosses into a new word; pos only ever advances by one)
  const long long nPix = (long long)a.H * a.W;
  if (ok && a.delta) {
    for (int d = 0; d < a.D && ok; d++) {
      T prevVal = 0;
      for (long long k = 0; k < nPix && ok; k++) {
        if (!a.allValidImage && !maskBit(a.bits, k)) continue;
        const int i = (int)(k / a.W), j = (int)(k - (long long)i * a.W);
        const int s = nextSymbol();
        if (!ok) break;
        T pred;
        if (j > 0 && (a.allValidImage || maskBit(a.bits, k - 1))) pred = prevVal;
        else if (i > 0 && (a.allValidImage || maskBit(a.bits, k - a.W))) pred = data[(k - a.W) * a.D + d];
        else pred = prevVal;
        const T val = (T)((T)(s - off) + pred);
        data[k * a.D + d] = val; prevVal = val;
      }
    }
  } else if (ok) {
    for (long long k = 0; k < nPix && ok; k++) {
      if (!a.allValidImage && !maskBit(a.bits, k)) continue;
      for (int d = 0; d < a.D; d++) { const int s = nextSymbol(); if (!ok) break; data[k * a.D + d] = (T)(s - off); }
    }
  }
  // the stream must hold the data words plus one read-ahead word (Lerc2.cpp:2583-2587)
  if (ok && 4 * (((pos & 31) ? 1 : 0) + (pos >> 5) + 1) > a.streamLen) ok = 0;
  if (!ok) atomicOr(a.status, DECF_BAD_STREAM);
}

}  // namespace lerc
#include "lerc_decode_fast.cuh"
#include "lerc_decode_stream.cuh"
#include "lerc_huffman_fast.cuh"
#include <cub/device/device_scan.cuh>
#include <atomic>
namespace lerc {

// =================================================================================================
// band orchestration (host)
namespace {


// k_dec_resolve: the region tables stay in global memory (default) or are staged in shared memory when they fit (experimental,
// LERC_B200_DEC_RESOLVE=smem)
inline void launchResolve(Context* ctx, const FastDecArgs& fa) {
  static const bool smemResolve = [] { const char* e = std::getenv("LERC_B200_DEC_RESOLVE"); return e && std::strcmp(e, "smem") == 0; }();
  const size_t bytes = (size_t)fa.nReg * FD_CAND * sizeof(FdEntry);
  if (smemResolve && bytes <= 160 * 1024) {
    static DeviceOnce attrSet;
    if (attrSet.need(ctx->device)) { cudaFuncSetAttribute(k_dec_resolve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attrSet.done(ctx->device); }
    LERC_LAUNCH(ctx, k_dec_resolve<true>, 1, 1024, bytes, fa, fa.nTx * fa.nTy);
  } else LERC_LAUNCH(ctx, k_dec_resolve<false>, 1, 1024, 0, fa, fa.nTx * fa.nTy);
}

// Single-kernel stream decoder (lerc_decode_stream.cuh): all-valid, nDepth == 1, 8x8 blocks, codec version >= 3.
// prefA / prefD: Fletcher partial sums of the band blob's bytes [14, streamPos) (host).  Returns 1 = decoded and the
// checksum is right, 0 = the kernel gave up (nothing is known about the blob: run the other decoders), -1 = checksum
// mismatch or CUDA error.
template <class T>
int decodeStreamFast(Context* ctx, const HeaderInfo& hd, DecodeBandArgs& ba, size_t streamPos, unsigned long long prefA, unsigned long long prefD) {
  const uint8_t* dBlob = ba.dBlob; void* dData = ba.dData;
  const size_t streamLen = (size_t)hd.blobSize - streamPos;
  if (streamLen == 0 || streamLen >= 0xfff00000ull || !std::isfinite(hd.zMax) || std::getenv("LERC_B200_NO_FAST")) return 0;
  static const bool trace = std::getenv("LERC_B200_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto us = [&]() { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count() * 1e-3; };
  cudaStream_t st = ctx->stream;
  constexpr int MAXSTRIPS = kMaxStrips;
  const int nChunks = (int)((streamLen + DS_CHUNK - 1) / DS_CHUNK);
  const size_t nGroups = ((size_t)nChunks + 31) / 32;
  const size_t stateBytes = 128 + ((size_t)nChunks * 2 + nGroups * 2) * 8;         // result | ticket counters | chunk and group states
  uint8_t* dState = (uint8_t*)ctx->arena.alloc(stateBytes);
  unsigned int* hStatus = (unsigned int*)ctx->pinnedAlloc(16);
  unsigned long long* hEnd = (unsigned long long*)ctx->pinnedAlloc(8 * MAXSTRIPS);
  if (!dState || !hStatus || !hEnd) return 0;
  cudaMemsetAsync(dState, 0, stateBytes, st);
  *hStatus = 0;
  StreamDecArgs sa;
  sa.stream = dBlob + streamPos; sa.streamLen = streamLen;
  sa.nRows = hd.nRows; sa.nCols = hd.nCols; sa.nTx = (hd.nCols + 7) / 8; sa.nTy = (hd.nRows + 7) / 8; sa.version = hd.version;
  sa.nTxMagic = sa.nTx >= 2 ? (uint32_t)((1ull << 32) / (unsigned)sa.nTx) + 1u : 0u;
  sa.invScale = 2 * hd.maxZError; sa.zMax = hd.zMax; sa.data = dData; sa.nChunks = nChunks;
  sa.res = (StreamDecResult*)dState;
  sa.exitState = (unsigned long long*)(dState + 128); sa.cntState = sa.exitState + nChunks;
  sa.groupState = sa.cntState + nChunks; sa.groupAcc = sa.groupState + nGroups;
  sa.regionOff = (long long)streamPos - 14; sa.regionLen = (long long)hd.blobSize - 14;
  sa.prefA = prefA % 65535ull; sa.prefD = prefD % 65535ull; sa.expectChecksum = hd.checksum; sa.haveChecksum = hd.version >= 3 ? 1 : 0;
  constexpr size_t smem = (size_t)DecStream<T>::SMEM;
  static std::atomic<unsigned long long> attrDone{0};
  const unsigned long long devBit = 1ull << (ctx->device & 63);
  if (!(attrDone.load(std::memory_order_relaxed) & devBit)) {
    if (!cudaOk(cudaFuncSetAttribute(k_decode_stream<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "decode stream smem")) return 0;
    attrDone.fetch_or(devBit, std::memory_order_relaxed);
  }
  // A host-resident call is pipelined: the stream travels to the device in strips (whole chunks) on one stream, every strip is decoded
  // as soon as it is there (exit chain and block counts carry over in the look-back state), and the block rows a strip has
  // completed go back to the caller's array on a third stream while the next strips are decoded.
  const size_t rowBytes = (size_t)hd.nCols * sizeof(T);
  int nStrips = 1, stripAt[kMaxStrips + 1] = {0, nChunks};
  if (ba.pendingBlob || ba.hOut) {
    // (the schedule follows the larger of the two transfers: the rows going back as a rule)
    nStrips = stripSchedule(std::max(ba.pendingBlob ? streamLen : 0, ba.hOut ? rowBytes * (size_t)hd.nRows : 0), nChunks, stripAt);
    if (nStrips > 1 && !ctx->pipeStreams()) { nStrips = 1; stripAt[1] = nChunks; }
  }
  const bool outPiped = nStrips > 1 && ba.hOut != nullptr;
  if (ba.pendingBlob) {
    uint8_t* dst = const_cast<uint8_t*>(dBlob);
    cudaStream_t cin = nStrips > 1 ? ctx->copyIn : st;
    size_t from = 0;
    for (int sI = 0; sI < nStrips; sI++) {
      const int ce = stripAt[sI + 1];
      const size_t to = sI == nStrips - 1 ? ba.pendingBlob : std::min(ba.pendingBlob, streamPos + (size_t)ce * DS_CHUNK + (size_t)DecStream<T>::LA + 32);
      if (to > from && !cudaOk(cudaMemcpyAsync(dst + from, ba.hBlob + from, to - from, cudaMemcpyHostToDevice, cin), "H2D strip")) return -1;
      from = std::max(from, to);
      if (nStrips > 1) cudaEventRecord(ctx->evStrip[0][sI], cin);
    }
    if (nStrips > 1) cudaStreamWaitEvent(st, ctx->evStrip[0][0], 0);
  }
  const bool inPiped = nStrips > 1 && ba.pendingBlob != 0;
  for (int sI = 0; sI < nStrips; sI++) {
    const int cb = stripAt[sI], ce = stripAt[sI + 1];
    sa.chunkBegin = cb; sa.ticket = (unsigned int*)(dState + 64) + sI;
    sa.hostStatus = outPiped ? hStatus : nullptr;
    sa.hostEnd = outPiped ? hEnd + sI : nullptr;                      // (blocks behind this strip, written by the kernel into mapped host memory)
    if (inPiped) cudaStreamWaitEvent(st, ctx->evStrip[0][sI], 0);
    { LaunchScope scope_(ctx, "k_decode_stream<T>"); k_decode_stream<T><<<(unsigned)(ce - cb), DS_THREADS, smem, st>>>(sa); ctx->kernelLaunches++; }
    if (outPiped) {
      cudaEventRecord(ctx->evStrip[1][sI], st);
    }
  }
  ba.pendingBlob = 0;                                                 // (the whole blob is in dBlob once the call's stream has passed the last strip's event)
  if (trace) std::fprintf(stderr, "[trace dec] %d strips enqueued at %.1f us\n", nStrips, us());
  if (outPiped) {
    int rowsSent = 0;
    bool ok = true;
    for (int sI = 0; sI < nStrips && ok; sI++) {
      ok = cudaOk(cudaEventSynchronize(ctx->evStrip[1][sI]), "strip sync");
      if (!ok) break;
      const unsigned long long blocks = hEnd[sI];
      const int rows = sI == nStrips - 1 ? hd.nRows : (int)std::min<unsigned long long>((unsigned long long)hd.nRows, blocks / (unsigned)sa.nTx * 8);
      if (trace) std::fprintf(stderr, "[trace dec] strip %d done at %.1f us, rows %d..%d\n", sI, us(), rowsSent, rows);
      if (rows > rowsSent) { cudaMemcpyAsync(ba.hOut + (size_t)rowsSent * rowBytes, (const uint8_t*)dData + (size_t)rowsSent * rowBytes, (size_t)(rows - rowsSent) * rowBytes, cudaMemcpyDeviceToHost, ctx->copyOut); rowsSent = rows; }
    }
    ok = cudaOk(cudaStreamSynchronize(ctx->copyOut), "copy out sync") && ok;
    if (trace) std::fprintf(stderr, "[trace dec] copies drained at %.1f us\n", us());
    if (!ok) return -1;
    if (!(*hStatus & DSF_REPORTED)) return -1;                        // (cannot happen: the last CTA of the last launch always reports)
    *hStatus &= ~DSF_REPORTED;
    ba.hostCopied = true;                                             // (void if the verdict below is a fallback: the caller copies again)
  } else if (!cudaOk(cudaMemcpyAsync(hStatus, &sa.res->status, 4, cudaMemcpyDeviceToHost, st), "D2H stream status") || !cudaOk(cudaStreamSynchronize(st), "sync")) return -1;
  if (!cudaOk(cudaGetLastError(), "k_decode_stream")) return -1;
  if (*hStatus && std::getenv("LERC_B200_VERBOSE")) std::fprintf(stderr, "[lerc_b200] stream decoder status %u\n", *hStatus);
  if (*hStatus & DSF_CHECKSUM) return -1;
  if (*hStatus & DSF_FALLBACK) { ba.hostCopied = false; return 0; }
  return 1;
}

// Block offsets of a masked band's stream with the stream decoder's boundary discovery (k_decode_stream<T, true>) + k_verify_offsets.
// 1 = dOff holds the serial parse's offsets, 0 = not established (other ways must find them), -1 = CUDA error.
template <class T>
int streamBlockOffsets(Context* ctx, const HeaderInfo& hd, const DecTileArgs& ta, int* dStatus) {
  const size_t streamLen = (size_t)ta.streamLen;
  if (streamLen == 0 || streamLen >= 0xfff00000ull || hd.nDepth != 1 || hd.microBlockSize != 8 || hd.version < 3 || std::getenv("LERC_B200_NO_FAST")) return 0;
  cudaStream_t st = ctx->stream;
  const int nChunks = (int)((streamLen + DS_CHUNK - 1) / DS_CHUNK);
  const size_t nGroups = ((size_t)nChunks + 31) / 32;
  const size_t stateBytes = 128 + ((size_t)nChunks * 2 + nGroups * 2) * 8;
  uint8_t* dState = (uint8_t*)ctx->arena.alloc(stateBytes);
  unsigned int* hStatus = (unsigned int*)ctx->pinnedAlloc(16);
  if (!dState || !hStatus) return 0;
  cudaMemsetAsync(dState, 0, stateBytes, st);
  cudaMemsetAsync(dStatus + 2, 0, 4, st);
  StreamDecArgs sa;
  std::memset(&sa, 0, sizeof sa);
  sa.stream = ta.stream; sa.streamLen = streamLen;
  sa.nRows = hd.nRows; sa.nCols = hd.nCols; sa.nTx = ta.nTx; sa.nTy = ta.nTy; sa.version = hd.version;
  sa.nTxMagic = sa.nTx >= 2 ? (uint32_t)((1ull << 32) / (unsigned)sa.nTx) + 1u : 0u;
  sa.invScale = 2 * hd.maxZError; sa.zMax = hd.zMax; sa.data = nullptr; sa.nChunks = nChunks;
  sa.res = (StreamDecResult*)dState;
  sa.exitState = (unsigned long long*)(dState + 128); sa.cntState = sa.exitState + nChunks;
  sa.groupState = sa.cntState + nChunks; sa.groupAcc = sa.groupState + nGroups;
  sa.haveChecksum = 0; sa.blockOff = ta.blockOff;
  sa.chunkBegin = 0; sa.ticket = (unsigned int*)(dState + 64);
  constexpr size_t smem = (size_t)DecStream<T>::SMEM_OFFS;
  static DeviceOnce attrDone;
  if (attrDone.need(ctx->device)) {
    if (!cudaOk(cudaFuncSetAttribute(k_decode_stream<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "stream offsets smem")) return 0;
    attrDone.done(ctx->device);
  }
  { LaunchScope scope_(ctx, "k_decode_stream<T, offsets>"); k_decode_stream<T, true><<<(unsigned)nChunks, DS_THREADS, smem, st>>>(sa); ctx->kernelLaunches++; }
  const int nBlocks = ta.nTx * ta.nTy;
  LERC_LAUNCH(ctx, k_verify_offsets, (unsigned)std::min((nBlocks + 127) / 128, 148 * 16), 128, 0, ta);
  if (!cudaOk(cudaMemcpyAsync(hStatus, &sa.res->status, 4, cudaMemcpyDeviceToHost, st), "D2H") ||
      !cudaOk(cudaMemcpyAsync(hStatus + 1, dStatus + 2, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return -1;
  if (!cudaOk(cudaGetLastError(), "k_decode_stream<offsets>")) return -1;
  if ((hStatus[0] || hStatus[1]) && std::getenv("LERC_B200_VERBOSE")) std::fprintf(stderr, "[lerc_b200] stream offsets status %u, verification %u\n", hStatus[0], hStatus[1]);
  return (hStatus[0] == 0 && hStatus[1] == 0) ? 1 : 0;
}

// Launches the speculative parallel decoder (lerc_decode_fast.cuh) on the micro-block stream.  Returns false when the
// stream's shape is outside what it handles (nothing launched).
template <class T>
bool launchDecodeFast(Context* ctx, const HeaderInfo& hd, const uint8_t* dStream, size_t streamLen, void* dData, int* dStatus,
                      const uint8_t* dBits = nullptr, uint32_t* dBlockOff = nullptr, FastDecArgs* faOut = nullptr) {
  // dBlockOff == nullptr: every pixel valid, decode the pixels too; else (masked raster) only the block offsets are produced
  const bool allValid = (long long)hd.numValidPixel == (long long)hd.nCols * hd.nRows;
  if (hd.nDepth != 1 || hd.microBlockSize != 8 || hd.version < 3 || (dBlockOff == nullptr) != allValid) return false;
  if (streamLen == 0 || streamLen >= 0xfff00000ull || std::getenv("LERC_B200_NO_FAST")) return false;
  const int nSub = (int)((streamLen + FD_SUB - 1) / FD_SUB);
  const int subPerReg = FD_REG;
  const int nReg = (nSub + subPerReg - 1) / subPerReg;
  const size_t smemB = fastDecodeBlocksSmem<T>(), smemW = fastDecodeWalkSmem<T>();
  const size_t szCand = (size_t)nSub * FD_CAND * sizeof(FdCand), szN = ((size_t)nSub + 255) & ~(size_t)255, szLens = (size_t)nSub * FD_CAND * FD_LENS,
               szSub = (size_t)nReg * FD_REG * FD_CAND * sizeof(FdEntry), szReg = (size_t)nReg * FD_CAND * sizeof(FdEntry), szEnt = ((size_t)(nReg + 1) * 8 + 255) & ~(size_t)255;
  uint8_t* scratch = (uint8_t*)ctx->arena.alloc(szLens + szCand + szSub + szReg + szEnt + szN + 256);
  if (!scratch) return false;
  FastDecArgs fa;
  fa.stream = dStream; fa.streamLen = streamLen;
  fa.nRows = hd.nRows; fa.nCols = hd.nCols; fa.nTx = (hd.nCols + 7) / 8; fa.nTy = (hd.nRows + 7) / 8; fa.dt = hd.dt; fa.version = hd.version;
  fa.invScale = 2 * hd.maxZError; fa.zMax = hd.zMax; fa.data = dData;
  fa.nSub = nSub; fa.subPerReg = subPerReg; fa.nReg = nReg;
  uint8_t* sp = scratch;
  fa.lens = sp; sp += szLens;
  fa.cand = (FdCand*)sp; sp += szCand;
  fa.subTab = (FdEntry*)sp; sp += szSub;
  fa.regTab = (FdEntry*)sp; sp += szReg;
  fa.regEntry = (uint32_t*)sp; sp += szEnt;
  fa.nCand = sp;
  fa.status = dStatus;
  fa.repairReg = -1; fa.repairPos = 0;
  static const bool closureDec = [] { const char* e = std::getenv("LERC_B200_DEC"); return e && std::strcmp(e, "closure") == 0; }();
  fa.firstOnly = closureDec ? 1 : 0;
  if (faOut) *faOut = fa;
  static DeviceOnce attrSet;
  if (attrSet.need(ctx->device)) {
    if (!cudaOk(cudaFuncSetAttribute(k_dec_blocks<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB), "smem attribute")) return false;
    if (!cudaOk(cudaFuncSetAttribute(k_dec_walk<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemW), "smem attribute")) return false;
    attrSet.done(ctx->device);
  }
  LERC_LAUNCH(ctx, k_dec_candidates<T>, (nSub + 7) / 8, 256, 0, fa);
  LERC_LAUNCH(ctx, k_dec_walk<T>, nReg, 256, smemW, fa);
  launchResolve(ctx, fa);
  if (dBlockOff) LERC_LAUNCH(ctx, k_dec_offsets<T>, nReg, FD_DWARPS * 32, 0, fa, dBits, dBlockOff);
  else LERC_LAUNCH(ctx, k_dec_blocks<T>, nReg, FD_DWARPS * 32, smemB, fa);
  return cudaOk(cudaGetLastError(), "launch fast decode");
}

// The speculative decoder gave up because the true chain entered a region at a position that was not among the region's kept
// candidates (a "candidate flood": long runs of 2-3-byte blocks make hundreds of byte positions parse as block headers that all
// merge into the true chain).  The true entry of that region is known from the resolve pass, so the region is
// repaired -- its first sub-chunk is walked from the true entry, closure and composition run again -- and the resolve pass is
// repeated, region after region, instead of handing the whole stream to the serial walk.  Returns true when the stream was
// decoded by the block-parallel kernels after all (status word cleared and checked by the caller as usual).
template <class T>
bool repairDecodeFast(Context* ctx, FastDecArgs fa, const uint8_t* dBits, uint32_t* dBlockOff, int* dStatus) {
  cudaStream_t st = ctx->stream;
  const size_t smemB = fastDecodeBlocksSmem<T>(), smemW = fastDecodeWalkSmem<T>();
  std::vector<uint32_t> ent(2 * ((size_t)fa.nReg + 1));
  int lastRepaired = -1;
  for (int iter = 0; iter < 64; iter++) {
    int hs = 0;
    if (!cudaOk(cudaMemcpyAsync(ent.data(), fa.regEntry, ent.size() * 4, cudaMemcpyDeviceToHost, st), "D2H region entries") ||
        !cudaOk(cudaMemcpyAsync(&hs, dStatus, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return false;
    if (iter > 0 && !(hs & DECF_FALLBACK)) {                       // the chain now runs through: decode the blocks
      if (dBlockOff) LERC_LAUNCH(ctx, k_dec_offsets<T>, fa.nReg, FD_DWARPS * 32, 0, fa, dBits, dBlockOff);
      else LERC_LAUNCH(ctx, k_dec_blocks<T>, fa.nReg, FD_DWARPS * 32, smemB, fa);
      return cudaOk(cudaGetLastError(), "launch fast decode (repaired)");
    }
    if (!(hs & 32)) return false;                                  // not the resolve pass that gave up
    int f = -1;
    for (int r = 1; r <= fa.nReg; r++) if (ent[2 * (size_t)r] == FD_DEAD) { f = r; break; }
    if (f < 1 || ent[2 * (size_t)(f - 1)] >= FD_DEAD - 1 || f - 1 == lastRepaired) return false;
    lastRepaired = f - 1;
    FastDecArgs ra = fa;
    ra.repairReg = f - 1; ra.repairPos = ent[2 * (size_t)(f - 1)];
    cudaMemsetAsync(dStatus, 0, 4, st);
    LERC_LAUNCH(ctx, k_dec_walk<T>, 1, 256, smemW, ra);
    launchResolve(ctx, fa);
  }
  return false;
}

// Parallel Huffman decode (lerc_huffman_fast.cuh) of an all-valid 8-bit band.  Returns 1 when the band was decoded,
// 0 when the serial kernel must run instead (unexpected table / no convergence / anomaly on the true chain), -1 on a CUDA error.
template <class T>
int decodeHuffmanFast(Context* ctx, const HuffmanTable& t, const uint8_t* dStream, size_t streamLen, int H, int W, int D, bool delta, void* dData) {
  if (std::getenv("LERC_B200_NO_FAST")) return 0;
  cudaStream_t st = ctx->stream;
  // ---- decode tables (host, 256 symbols)
  HuffFastTables* hT = (HuffFastTables*)ctx->pinnedAlloc(sizeof(HuffFastTables));
  if (!hT) return 0;
  std::memset(hT, 0, sizeof *hT);
  hT->minLen = 33; hT->maxLen = 0;
  std::vector<std::pair<uint32_t, int>> byLen[33];
  for (int s = 0; s < 256; s++) if (t.len[s]) {
    if (t.len[s] > 32) return 0;
    byLen[t.len[s]].push_back({t.code[s], s});
    hT->minLen = std::min<int>(hT->minLen, t.len[s]); hT->maxLen = std::max<int>(hT->maxLen, t.len[s]);
  }
  if (hT->maxLen == 0) return 0;
  int nSyms = 0;
  for (int len = 1; len <= 32; len++) {
    auto& v = byLen[len];
    if (v.empty()) continue;
    std::sort(v.begin(), v.end());
    hT->first[len] = v[0].first; hT->count[len] = (uint32_t)v.size(); hT->offset[len] = (uint32_t)nSyms;
    for (size_t i = 0; i < v.size(); i++) {
      if (v[i].first != v[0].first + i || (len < 32 && (v[i].first >> len))) return 0;       // codes of one length must be consecutive (Huffman.cpp:541-572)
      hT->syms[nSyms++] = (uint8_t)v[i].second;
      if (len <= 12) for (uint32_t k = v[i].first << (12 - len); k < ((v[i].first + 1) << (12 - len)); k++) {
        if (hT->lut[k]) return 0;                                                                // not prefix free
        hT->lut[k] = (uint16_t)((len << 8) | v[i].second);
      }
    }
  }
  {  // the codes must be prefix free (the reference's tree build and the serial kernel's trie reject the table otherwise)
    struct LC { uint32_t aligned; int len; };
    LC all[256]; int nAll = 0;
    for (int sIdx = 0; sIdx < 256; sIdx++) if (t.len[sIdx]) { all[nAll].len = t.len[sIdx]; all[nAll].aligned = t.len[sIdx] == 32 ? t.code[sIdx] : (t.code[sIdx] << (32 - t.len[sIdx])); nAll++; }
    for (int i = 0; i < nAll; i++)
      for (int j = 0; j < nAll; j++)
        if (i != j && all[i].len <= all[j].len && (all[j].aligned >> (32 - all[i].len)) == (all[i].aligned >> (32 - all[i].len))) return 0;
  }
  const unsigned long long nBits = (unsigned long long)(streamLen / 4) * 32;
  if (nBits < 64) return 0;
  const unsigned long long nSym = (unsigned long long)H * W * D;
  const int nChunks = (int)((nBits + HF_CHUNK - 1) / HF_CHUNK);
  HuffFastTables* dT = (HuffFastTables*)ctx->arena.alloc(sizeof(HuffFastTables));
  unsigned long long* dArr = (unsigned long long*)ctx->arena.alloc((size_t)nChunks * 8 * 3);
  uint32_t* dCount = (uint32_t*)ctx->arena.alloc((size_t)(nChunks + 1) * 4);
  uint32_t* dBase = (uint32_t*)ctx->arena.alloc((size_t)(nChunks + 1) * 4);
  int* dFlags = (int*)ctx->arena.alloc(32);
  int* hFlags = (int*)ctx->pinnedAlloc(32);
  if (!dT || !dArr || !dCount || !dBase || !dFlags || !hFlags) return 0;
  cudaMemcpyAsync(dT, hT, sizeof(HuffFastTables), cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(dFlags, 0, 32, st);
  cudaMemsetAsync(dCount + nChunks, 0, 4, st);
  HuffFastArgs ha;
  ha.stream = dStream; ha.nBits = nBits; ha.tab = dT; ha.nSym = nSym; ha.nChunks = nChunks;
  ha.startA = dArr; ha.endA = dArr + nChunks; ha.endB = dArr + 2 * (size_t)nChunks; ha.count = dCount; ha.changed = dFlags; ha.bad = dFlags + 1; ha.endBit = (unsigned long long*)(dFlags + 4);
  const int grid = (nChunks + 255) / 256;
  int iter = 0;
  for (;; iter++) {
    if (iter >= 64) return 0;                                        // does not synchronise: let the serial decoder handle it
    if (iter > 0) cudaMemsetAsync(dFlags, 0, 4, st);
    LERC_LAUNCH(ctx, k_huff_chunks, grid, 256, 0, ha, iter);
    if (iter == 0) continue;
    if (!cudaOk(cudaMemcpyAsync(hFlags, dFlags, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return -1;
    if (!hFlags[0]) break;
  }
  const unsigned long long* dEndFinal = (iter & 1) ? ha.endB : ha.endA;   // written by the last iteration
  {
    LaunchScope scope(ctx, "cub::ExclusiveSum<u32>");
    size_t tmpBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, dCount, dBase, nChunks + 1, st);
    void* tmp = ctx->arena.alloc(tmpBytes ? tmpBytes : 16);
    if (!tmp) return 0;
    cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, dCount, dBase, nChunks + 1, st);
    ctx->kernelLaunches += 2;
  }
  uint32_t total = 0;
  if (!cudaOk(cudaMemcpyAsync(hFlags, dBase + nChunks, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return -1;
  total = (uint32_t)hFlags[0];
  if ((unsigned long long)total < nSym || nSym > 0xffffffffull) return 0;     // stream holds fewer symbols than the band: the serial decoder reports the error
  const int off = PixelTraits<T>::code == DT_Char ? 128 : 0;
  uint8_t* dPlanes = nullptr;
  if (delta) { dPlanes = (uint8_t*)ctx->arena.alloc((size_t)nSym + 16); if (!dPlanes) return 0; }
  LERC_LAUNCH(ctx, k_huff_emit<T>, grid, 256, 0, ha, dEndFinal, dBase, delta ? dPlanes : (uint8_t*)dData);
  if (!delta) {
    if (off) LERC_LAUNCH(ctx, k_huff_values<T>, 148 * 8, 256, 0, (uint8_t*)dData, nSym, off);
  } else {
    uint8_t* dCol0 = (uint8_t*)ctx->arena.alloc((size_t)H * D + 16);
    if (!dCol0) return 0;
    LERC_LAUNCH(ctx, k_huff_col0, D, 256, 0, dPlanes, H, W, D, off, dCol0);
    const bool vec = (D == 1 || D == 3) && W % 16 == 0 && (((uintptr_t)dPlanes | (uintptr_t)dData) & 15) == 0;
    if (vec && D == 3) LERC_LAUNCH(ctx, k_huff_rows_vec<3>, H, 256, 0, dPlanes, dCol0, H, W, off, (uint8_t*)dData);
    else if (vec) LERC_LAUNCH(ctx, k_huff_rows_vec<1>, H, 256, 0, dPlanes, dCol0, H, W, off, (uint8_t*)dData);
    else LERC_LAUNCH(ctx, k_huff_rows<T>, H, 256, 0, dPlanes, dCol0, H, W, D, off, (T*)dData);
  }
  if (!cudaOk(cudaMemcpyAsync(hFlags, dFlags, 32, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return -1;
  if (hFlags[1]) return 0;
  {  // the stream must hold the data words plus the read-ahead word (Lerc2.cpp:2583-2587)
    unsigned long long endBit; std::memcpy(&endBit, hFlags + 4, 8);
    if (endBit == 0 || 4 * (((endBit & 31) ? 1 : 0) + (size_t)(endBit >> 5) + 1) > streamLen) return 0;
  }
  globalStats().fastPathDecodes++;
  return 1;
}

}  // namespace
}  // namespace lerc
#include "lerc_fpl_decode.cuh"
namespace lerc {
namespace {

template <class T>
ErrCode decodeBandT(Context* ctx, DecodeBandArgs& a, BandMaskState& ms) {
  const HeaderInfo& hd = a.hd;
  cudaStream_t st = ctx->stream;
  const long long nPix = (long long)hd.nCols * hd.nRows;
  const size_t nBits = (size_t)((nPix + 7) >> 3);
  const int nDepth = hd.nDepth;
  if (a.avail < (size_t)hd.blobSize) return Failed;
  const uint8_t* blob = a.dBlob;
  // reads of this band's blob on the host go through the caller's ByteSource (host pointer, or device pointer + cache)
  struct BandView { const ByteSource* s; size_t off, size;
                    bool fetch(size_t o, size_t len, void* dst) const { return o <= size && len <= size - o && s->fetch(off + o, len, dst); } };
  ByteSource own;
  if (!a.src) { own.base = a.hBlob ? a.hBlob : a.dBlob; own.size = (size_t)hd.blobSize; own.onDevice = a.hBlob == nullptr; }
  const BandView src{a.src ? a.src : &own, a.src ? a.srcOff : 0, (size_t)hd.blobSize};

  // status words of the kernels other than the stream decoder (allocated and cleared on first use: the stream decoder has its own)
  int* dStatus = nullptr;
  auto needStatus = [&]() -> bool {
    if (dStatus) return true;
    dStatus = (int*)ctx->arena.alloc(16);
    if (!dStatus) return false;
    cudaMemsetAsync(dStatus, 0, 16, st);
    return true;
  };
  // the bit mask of an all-valid / all-invalid band is only written when something reads it
  auto bitsNow = [&]() { if (ms.pendingFill >= 0) { cudaMemsetAsync(ms.dBits, ms.pendingFill, nBits, st); ms.pendingFill = -1; } };

  // a host blob that is not on the device yet: only the stream decoder below copies it itself (in strips); every other path wants it whole
  // (they all start the checksum kernel first, which therefore stages the blob)
  const bool mayFast = nDepth == 1 && hd.numValidPixel == nPix && hd.microBlockSize == 8 && hd.version >= 3 && hd.zMin != hd.zMax;
  auto stageBlob = [&]() {
    if (a.pendingBlob) cudaMemcpyAsync(const_cast<uint8_t*>(a.dBlob), a.hBlob, a.pendingBlob, cudaMemcpyHostToDevice, st);
    a.pendingBlob = 0;
  };
  if (!mayFast) stageBlob();

  // checksum (Lerc2.cpp:592-601); the verdict is read together with the other status bits at the end.  The single-kernel stream
  // decoder sums the blob itself, so the launch waits until it is known whether that decoder applies.
  if (hd.version >= 3 && hd.blobSize < 14) return Failed;
  bool checksumLaunched = false;
  auto launchChecksum = [&]() -> bool {
    stageBlob();
    if (hd.version < 3 || checksumLaunched) return true;
    unsigned long long* dAcc = (unsigned long long*)ctx->arena.alloc(16);
    if (!dAcc) return false;
    if (!needStatus()) return false;
    ctx->forkSide();                                   // the checksum only reads the blob: it runs beside the decode kernels
    cudaMemsetAsync(dAcc, 0, 16, ctx->stream);
    launchFletcher(ctx, blob + 14, (long long)hd.blobSize - 14, dAcc, nullptr, hd.checksum, dStatus);
    ctx->backToMain();
    checksumLaunched = true;
    return true;
  };

  // mask (Lerc2.cpp:961-1008)
  size_t pos = (size_t)headerBytes(hd.version);
  int32_t nm = 0;
  if (!src.fetch(pos, 4, &nm)) return Failed;
  pos += 4;
  if (nm < 0 || ((hd.numValidPixel == 0 || hd.numValidPixel == nPix) && nm != 0)) return Failed;
  if (hd.numValidPixel == 0) { ms.pendingFill = 0; ms.havePrev = true; }
  else if (hd.numValidPixel == nPix) { ms.pendingFill = 0xff; ms.havePrev = true; }
  else if (nm > 0) {
    if (pos + (size_t)nm > (size_t)hd.blobSize || !needStatus()) return Failed;
    ms.pendingFill = -1;
    int* dRleOk = dStatus + 1;
    launchRleDecode(ctx, blob + pos, (long long)(hd.blobSize - pos), ms.dBits, (long long)nBits, dRleOk);
    pos += (size_t)nm; ms.havePrev = true;
    ms.numValid = -1;    // (the RLE status is checked below)
  } else if (!ms.havePrev) return Failed;          // "same as previous band" without a previous band (Lerc2.cpp:1002)
  else bitsNow();                                  // the previous band's mask is this band's
  const bool rleUsed = nm > 0;
  if (a.dValidBytes) { bitsNow(); launchBitsToBytes(ctx, ms.dBits, nPix, a.dValidBytes); }           // Lerc.cpp:481, :979-995

  // Lerc2.cpp:609 zero-fills the output; when every pixel is valid and coded by the micro-block stream each one is
  // overwritten, so the fill is deferred until the fused decoder is known not to apply.
  bool zeroFilled = false;
  auto zeroFill = [&]() { if (!zeroFilled) { cudaMemsetAsync(a.dData, 0, (size_t)nPix * nDepth * sizeof(T), st); zeroFilled = true; } };
  if (!mayFast) { bitsNow(); zeroFill(); if (!needStatus() || !launchChecksum()) return Failed; }

  auto finish = [&]() -> ErrCode {
    int hStatus[2] = {0, 1};
    if (!needStatus() || !launchChecksum()) return Failed;              // (paths that did not launch it earlier)
    ctx->joinSide();
    if (!cudaOk(cudaMemcpyAsync(hStatus, dStatus, 8, cudaMemcpyDeviceToHost, st), "D2H status")) return Failed;
    if (!cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
    if (hStatus[0] != 0) return Failed;
    if (rleUsed && hStatus[1] != 1) return Failed;
    return Ok;
  };
  const int fillGrid = (int)std::min<long long>((nPix * nDepth + 255) / 256, 148 * 32);

  if (hd.numValidPixel == 0) return finish();
  if (hd.zMin == hd.zMax) {
    bitsNow();
    LERC_LAUNCH(ctx, k_fill_const<T>, fillGrid, 256, 0, (T*)a.dData, ms.dBits, nPix, nDepth, (double)(T)hd.zMin, (const double*)nullptr);
    return finish();
  }
  double* dZMax = nullptr;
  if (hd.version >= 4) {                                                                // Lerc2.cpp:2643-2677
    const size_t len = (size_t)nDepth * sizeof(T);
    if (pos + 2 * len > (size_t)hd.blobSize) return Failed;
    std::vector<T> r(2 * (size_t)nDepth);
    if (!src.fetch(pos, 2 * len, r.data())) return Failed;
    pos += 2 * len;
    std::vector<double> zr(2 * (size_t)nDepth);
    for (int i = 0; i < 2 * nDepth; i++) zr[i] = (double)r[i];
    const bool allConst = 0 == std::memcmp(zr.data(), zr.data() + nDepth, sizeof(double) * nDepth);
    double* dRanges = nullptr;
    if (allConst || nDepth > 1) {                        // the kernels read the ranges only then
      dRanges = (double*)ctx->arena.alloc(16 * (size_t)nDepth);
      double* hRanges = (double*)ctx->pinnedAlloc(16 * (size_t)nDepth);
      if (!dRanges || !hRanges) return Failed;
      std::memcpy(hRanges, zr.data(), 16 * (size_t)nDepth);
      cudaMemcpyAsync(dRanges, hRanges, 16 * (size_t)nDepth, cudaMemcpyHostToDevice, st);
    }
    if (allConst) {   // every depth constant
      bitsNow();
      LERC_LAUNCH(ctx, k_fill_const<T>, fillGrid, 256, 0, (T*)a.dData, ms.dBits, nPix, nDepth, 0.0, (const double*)dRanges);
      return finish();
    }
    if (nDepth > 1) dZMax = dRanges + nDepth;
  }
  if (pos + 1 > (size_t)hd.blobSize) return Failed;
  uint8_t flags[2] = {0, 0};
  if (!src.fetch(pos, std::min<size_t>(2, (size_t)hd.blobSize - pos), flags)) return Failed;
  pos += 1;
  if (flags[0]) {                                                                       // one sweep
    bitsNow(); zeroFill();
    if (!needStatus() || !launchChecksum()) return Failed;
    const size_t len = (size_t)nDepth * sizeof(T);
    // numValidPixel of the header is trusted here only after comparing with the mask popcount on the device path below
    if (hd.numValidPixel == nPix) {
      if (pos + len * (size_t)nPix > (size_t)hd.blobSize) return Failed;
      cudaMemcpyAsync(a.dData, blob + pos, len * (size_t)nPix, cudaMemcpyDeviceToDevice, st);
    } else {
      const int nChunks = (int)((nPix + 1023) >> 10);
      uint32_t* cnt = (uint32_t*)ctx->arena.alloc(4 * (size_t)(nChunks + 1));
      uint32_t* base = (uint32_t*)ctx->arena.alloc(4 * (size_t)(nChunks + 1));
      if (!cnt || !base) return Failed;
      cudaMemsetAsync(cnt, 0, 4 * (size_t)(nChunks + 1), st);
      launchChunkValidCounts(ctx, ms.dBits, nPix, nChunks, cnt);
      exclusiveScanU32(ctx, cnt, base, (size_t)nChunks);
      uint32_t total = 0;
      if (!cudaOk(cudaMemcpyAsync(&total, base + nChunks, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
      if (pos + len * (size_t)total > (size_t)hd.blobSize) return Failed;             // Lerc2.cpp:1385
      LERC_LAUNCH(ctx, k_one_sweep_scatter<T>, std::min((nChunks + 7) / 8, 148 * 8), 256, 0, (T*)a.dData, ms.dBits, base, nPix, nDepth, blob + pos);
    }
    return finish();
  }
  if (hd.tryHuffmanInt() || hd.tryHuffmanFlt()) {
    if (pos + 1 > (size_t)hd.blobSize) return Failed;
    const int mode = flags[1];
    pos += 1;
    if (mode > 3 || (mode > 2 && hd.version < 6) || (mode > 1 && hd.version < 4)) return Failed;
    if (mode != IEM_Tiling) {
      bitsNow(); zeroFill();
      if (!needStatus() || !launchChecksum()) return Failed;
      if constexpr (sizeof(T) == 1) {
        if (!(hd.tryHuffmanInt() && (mode == IEM_DeltaHuffman || (hd.version >= 4 && mode == IEM_Huffman)))) return Failed;
        std::vector<uint8_t> tb(std::min<size_t>(2048, (size_t)hd.blobSize - pos));
        if (!src.fetch(pos, tb.size(), tb.data())) return Failed;
        HuffmanTable t;
        const size_t used = t.read(tb.data(), tb.size(), hd.version);
        if (!used) return Failed;
        pos += used;
        HuffDecArgs ha;
        ha.stream = blob + pos; ha.streamLen = (unsigned long long)((size_t)hd.blobSize - pos);
        ha.bits = ms.dBits; ha.H = hd.nRows; ha.W = hd.nCols; ha.D = nDepth; ha.delta = mode == IEM_DeltaHuffman;
        ha.allValidImage = hd.numValidPixel == nPix;
        std::memcpy(ha.len, t.len, sizeof ha.len); std::memcpy(ha.code, t.code, sizeof ha.code);
        ha.data = a.dData; ha.status = dStatus;
        if (ha.allValidImage) {                         // parallel decoder first (lerc_huffman_fast.cuh)
          const int rc = decodeHuffmanFast<T>(ctx, t, ha.stream, (size_t)ha.streamLen, hd.nRows, hd.nCols, nDepth, ha.delta != 0, a.dData);
          if (rc < 0) return Failed;
          if (rc > 0) return finish();
        }
        LERC_LAUNCH(ctx, k_huffman_decode_seq<T>, 1, 32, 0, ha);
        return finish();
      } else if constexpr (PixelTraits<T>::isFloat) {
        // the reference's lossless float codec (fpl_*.cpp; lerc_fpl_decode.cuh): codes the whole array, invalid pixels included
        if (!(hd.tryHuffmanFlt() && mode == IEM_DeltaDeltaHuffman)) return Failed;                                   // Lerc2.cpp:674-680
        const ErrCode fe = decodeFpl<T>(ctx, [&](size_t o, size_t len, void* dst) { return src.fetch(o, len, dst); }, blob, pos, (size_t)hd.blobSize,
                                        hd.nCols, hd.nRows, nDepth, a.dData, dStatus);
        if (fe != Ok) return fe;
        return finish();
      } else return Failed;
    }
  }
  // micro-block stream: the single-kernel stream decoder first (lerc_decode_stream.cuh) ...
  if (mayFast) {
    unsigned long long pA = 0, pD = 0;
    uint8_t pre[256];
    if (pos > 14 && pos - 14 <= sizeof pre && src.fetch(14, pos - 14, pre)) {
      fletcherHostPartial(pre, 0, (long long)pos - 14, pA, pD);
      const int rc = decodeStreamFast<T>(ctx, hd, a, pos, pA, pD);
      if (rc < 0) return Failed;
      if (rc > 0) { globalStats().fastPathDecodes++; return Ok; }
    }
    stageBlob();
    bitsNow();
    if (!needStatus() || !launchChecksum()) return Failed;
  }
  // ... then the multi-kernel speculative decoder (lerc_decode_fast.cuh: keeps up to 16 entry candidates per sub-chunk, repairs floods)
  FastDecArgs fdArgs;
  if (mayFast && launchDecodeFast<T>(ctx, hd, blob + pos, (size_t)hd.blobSize - pos, a.dData, dStatus, nullptr, nullptr, &fdArgs)) {
    int hs = 0;
    ctx->joinSide();
    if (!cudaOk(cudaMemcpyAsync(&hs, dStatus, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
    if (hs && std::getenv("LERC_B200_VERBOSE")) std::fprintf(stderr, "[lerc_b200] fused decoder status %d\n", hs);
    if (!(hs & DECF_FALLBACK)) { if (hs == 0) globalStats().fastPathDecodes++; return (hs & 7) == 0 ? Ok : Failed; }
    if (hs & 7) return Failed;                                  // e.g. checksum mismatch (bits above DECF_FALLBACK say why the fused decoder gave up)
    if ((hs & 32) && repairDecodeFast<T>(ctx, fdArgs, nullptr, nullptr, dStatus)) {            // candidate flood: repair the regions the chain could not enter
      if (!cudaOk(cudaMemcpyAsync(&hs, dStatus, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
      if (hs && std::getenv("LERC_B200_VERBOSE")) std::fprintf(stderr, "[lerc_b200] fused decoder status after repair %d\n", hs);
      if (hs == 0) { globalStats().fastPathDecodes++; return Ok; }
      if (hs & 7) return Failed;
    }
    cudaMemsetAsync(dStatus, 0, 4, st);
  }
  zeroFill();
  DecTileArgs ta; std::memset(&ta, 0, sizeof ta);
  ta.stream = blob + pos; ta.streamLen = (unsigned long long)((size_t)hd.blobSize - pos);
  ta.bits = ms.dBits; ta.nRows = hd.nRows; ta.nCols = hd.nCols; ta.nDepth = nDepth; ta.mb = hd.microBlockSize;
  ta.nTx = (hd.nCols + ta.mb - 1) / ta.mb; ta.nTy = (hd.nRows + ta.mb - 1) / ta.mb;
  ta.dt = hd.dt; ta.version = hd.version; ta.allValidImage = hd.numValidPixel == nPix;
  ta.maxZErr = hd.maxZError; ta.zMaxHdr = hd.zMax; ta.zMaxVec = dZMax;
  const size_t nBlocks = (size_t)ta.nTx * ta.nTy;
  ta.blockOff = (uint32_t*)ctx->arena.alloc(4 * (nBlocks + 1));
  ta.data = a.dData; ta.status = dStatus;
  if (!ta.blockOff) return Failed;
  // block boundaries: speculative parallel discovery for masked nDepth == 1 rasters (lerc_decode_fast.cuh), else / on
  // any inconsistency the exact serial walk
  bool haveOffsets = false;
  if (hd.numValidPixel != nPix && !ta.allValidImage && nDepth == 1) {
    const int rc = streamBlockOffsets<T>(ctx, hd, ta, dStatus);
    if (rc < 0) return Failed;
    if (rc > 0) { haveOffsets = true; globalStats().fastPathDecodes++; }
  }
  if (!haveOffsets && hd.numValidPixel != nPix && !ta.allValidImage && nDepth == 1 &&
      launchDecodeFast<T>(ctx, hd, ta.stream, (size_t)ta.streamLen, nullptr, dStatus, ms.dBits, ta.blockOff, &fdArgs)) {
    int hs = 0;
    ctx->joinSide();
    if (!cudaOk(cudaMemcpyAsync(&hs, dStatus, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
    if (hs && std::getenv("LERC_B200_VERBOSE")) std::fprintf(stderr, "[lerc_b200] speculative block offsets status %d\n", hs);
    if (hs & 7) return Failed;
    if ((hs & DECF_FALLBACK) && (hs & 32) && repairDecodeFast<T>(ctx, fdArgs, ms.dBits, ta.blockOff, dStatus)) {
      if (!cudaOk(cudaMemcpyAsync(&hs, dStatus, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
      if (hs & 7) return Failed;
    }
    if (!(hs & DECF_FALLBACK)) { haveOffsets = true; globalStats().fastPathDecodes++; }
    else cudaMemsetAsync(dStatus, 0, 4, st);
  }
  if (!haveOffsets) {
    LERC_LAUNCH(ctx, k_walk_units, 1, 128, 0, ta);
    // a malformed chain must not reach the unpack kernel (its offsets would be garbage)
    int hs = 0;
    ctx->joinSide();
    if (!cudaOk(cudaMemcpyAsync(&hs, dStatus, 4, cudaMemcpyDeviceToHost, st), "D2H") || !cudaOk(cudaStreamSynchronize(st), "sync")) return Failed;
    if (hs & DECF_BAD_STREAM) return Failed;
  }
  const int warps = 8;
  int grid = (int)std::min<size_t>((nBlocks + warps - 1) / warps, (size_t)148 * 64);
  LERC_LAUNCH(ctx, k_tiles_decode<T>, grid, warps * 32, 0, ta);
  return finish();
}

}  // namespace

ErrCode decodeBand(Context* ctx, DecodeBandArgs& a, BandMaskState& ms) {
  if (a.hd.dt != a.dt) return Failed;     // deviation: the reference reinterprets (DESIGN.md "Deviations")
  switch (a.dt) {
    case DT_Char:   return decodeBandT<int8_t>(ctx, a, ms);
    case DT_Byte:   return decodeBandT<uint8_t>(ctx, a, ms);
    case DT_Short:  return decodeBandT<int16_t>(ctx, a, ms);
    case DT_UShort: return decodeBandT<uint16_t>(ctx, a, ms);
    case DT_Int:    return decodeBandT<int32_t>(ctx, a, ms);
    case DT_UInt:   return decodeBandT<uint32_t>(ctx, a, ms);
    case DT_Float:  return decodeBandT<float>(ctx, a, ms);
    case DT_Double: return decodeBandT<double>(ctx, a, ms);
    default: return WrongParam;
  }
}

}  // namespace lerc

#include "lerc_tiles_decode.cuh"
