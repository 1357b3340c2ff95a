// lerc_internal.h -- declarations shared by the host side and the CUDA translation units of lerc_b200.
//
// Layering:
//   lerc_capi.cpp      C ABI (include/Lerc_c_api.h, include/lerc_b200.h): argument checks, pointer
//                      classification (host / pinned / device), band loop
//   lerc_format.cpp    pure host: Lerc2 header read/write, multi-band header walk, Huffman code
//                      table construction and (de)serialisation
//   lerc_encode.cu     band encoder: statistics, micro-block sizing/packing, Huffman, one-sweep
//   lerc_decode.cu     band decoder: block boundary discovery, unpack/dequantise, Huffman, one-sweep
//   lerc_mask.cu       validity mask kernels (byte<->bit, RLE), Fletcher-32, small utilities
//
// Reference citations are relative to /root/reference/src/LercLib.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstdint>
#include <cstddef>
#include <vector>
#include <cuda_runtime.h>

namespace lerc {

enum ErrCode : unsigned { Ok = 0, Failed, WrongParam, BufferTooSmall, NaNFound, HasNoData, DimensionsTooLarge };
enum DataType : int { DT_Char = 0, DT_Byte, DT_Short, DT_UShort, DT_Int, DT_UInt, DT_Float, DT_Double, DT_Undefined };
enum ImageEncodeMode : int { IEM_Tiling = 0, IEM_DeltaHuffman = 1, IEM_Huffman = 2, IEM_DeltaDeltaHuffman = 3 };

inline int typeSize(int dt) { static const int s[8] = {1, 1, 2, 2, 4, 4, 4, 8}; return (dt >= 0 && dt < 8) ? s[dt] : 0; }

// ---------------------------------------------------------------------------------------------
// Lerc2 band header (Lerc2.h:102-131, Lerc2.cpp:710-917)
struct HeaderInfo {
  int version = 6;
  uint32_t checksum = 0;
  int nRows = 0, nCols = 0, nDepth = 1, numValidPixel = 0, microBlockSize = 8, blobSize = 0, dt = DT_Undefined, nBlobsMore = 0;
  uint8_t bPassNoDataValues = 0, bIsInt = 0, bReserved3 = 0, bReserved4 = 0;
  double maxZError = 0, zMin = 0, zMax = 0, noDataVal = 0, noDataValOrig = 0;

  bool tryHuffmanInt() const { return version >= 2 && (dt == DT_Byte || dt == DT_Char) && maxZError == 0.5; }
  bool tryHuffmanFlt() const { return version >= 6 && (dt == DT_Float || dt == DT_Double) && maxZError == 0; }
};

int headerBytes(int version);                                        // 58 / 62 / 66 / 90
void writeHeader(uint8_t* dst, const HeaderInfo& hd);                // checksum slot written as hd.checksum
bool readHeader(const uint8_t* src, size_t avail, HeaderInfo& hd);   // incl. the guards of Lerc2.cpp:877-911

// A blob that may live in host or device memory; fetch() brings small pieces to the host.
struct ByteSource {
  const uint8_t* base = nullptr;
  size_t size = 0;
  bool onDevice = false;
  cudaStream_t stream = nullptr;   // device blobs: the stream the call runs on (fetches are ordered behind the work that produced the blob); nullptr = legacy default stream
  bool fetch(size_t off, size_t len, void* dst) const;
  // device blobs: a small host copy of the bytes around the last fetch, so that the handful of header / mask length /
  // range / flag reads of one band cost one device-to-host copy (valid for the duration of one API call only)
  mutable uint8_t cache[512];
  mutable size_t cacheOff = 0, cacheLen = 0;
};

struct BlobInfo {     // Lerc.h:99-116
  int version = 0, nDepth = 0, nCols = 0, nRows = 0, numValidPixel = 0, nBands = 0, nMasks = 0, dt = 0, nUsesNoDataValue = 0;
  uint32_t blobSize = 0;
  double zMin = 0, zMax = 0, maxZError = 0;
};
// Lerc::GetLercInfo (Lerc.cpp:92-182), Lerc2 blobs only.
ErrCode getBlobInfo(const ByteSource& src, BlobInfo& info, double* mins, double* maxs, size_t nElem);

// ---------------------------------------------------------------------------------------------
// canonical Huffman code table for the 8-bit path (Huffman.cpp)
struct HuffmanTable {
  uint16_t len[256];
  uint32_t code[256];
  bool buildFromHistogram(const int* histo);              // Huffman.cpp:35-81 + :541-572
  bool range(int& i0, int& i1, int& maxLen) const;        // Huffman.cpp:383-438
  bool tableBytes(int& nBytes) const;                     // Huffman.cpp:357-379
  bool totalBytes(const int* histo, int& nBytes) const;   // Huffman.cpp:85-111
  size_t write(uint8_t* dst, int version = 6) const;      // Huffman.cpp:126-166; dst zero-filled; returns bytes (version 2: MSB-first code lengths)
  size_t read(const uint8_t* src, size_t avail, int version = 6);   // Huffman.cpp:170-234; returns bytes consumed or 0
};

// ---------------------------------------------------------------------------------------------
// CUDA context: one stream + a growable device arena + pinned staging, pooled per concurrent call.
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, used = 0;
  std::vector<uint8_t*> retired;   // blocks replaced by a bigger one during this call; freed at release
  void* alloc(size_t bytes, size_t align = 256);
  void reset();
};

struct Context {
  int device = 0;
  cudaStream_t stream = nullptr;
  // side stream for work that is independent of the main kernel chain (checksum of the blob on decode, the row-0
  // maxZError test on encode): forkSide() makes it wait for everything issued on `stream` so far and switches `stream`
  // to it; joinSide() switches back and makes `stream` wait for the side work.
  cudaStream_t side = nullptr, mainSaved = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  // host <-> device pipelines of the single-pass paths (strips of a host raster / blob move while earlier strips are coded): one stream
  // per direction and a few events without timing, created on first use
  cudaStream_t copyIn = nullptr, copyOut = nullptr;
  cudaEvent_t evStrip[2][16] = {};
  bool pipeStreams();
  bool sidePending = false;
  bool drainOnRelease = true;   // releaseContext synchronises the call's stream unless the call says its stream is idle / its results are stream-ordered
  void forkSide();
  void backToMain();
  void joinSide();
  Arena arena;              // device scratch, bump-allocated per API call
  uint8_t* pinned = nullptr;   // pinned host staging for small control transfers
  size_t pinnedCap = 0;
  uint8_t* pinnedBig = nullptr; size_t pinnedBigCap = 0;   // staging for pageable host payloads
  uint64_t kernelLaunches = 0; // counted for lerc_b200_stats()
  void* pinnedAlloc(size_t bytes);   // returns a region of `pinned` (bump, reset per call)
  size_t pinnedUsed = 0;
  // optional per-kernel timing (lerc_b200_profile): event pairs recorded around every launch
  struct ProfRec { const char* name; cudaEvent_t a, b; };
  std::vector<ProfRec> profRecs;
  std::vector<cudaEvent_t> eventPool;
  cudaEvent_t takeEvent();
};

extern bool gProfileEnabled;
// RAII: records an event pair around the launches issued while it is alive (only when profiling is on)
struct LaunchScope {
  Context* ctx; cudaEvent_t a = nullptr; const char* name;
  LaunchScope(Context* c, const char* n) : ctx(c), name(n) { if (gProfileEnabled) { a = ctx->takeEvent(); cudaEventRecord(a, ctx->stream); } }
  ~LaunchScope() { if (a) { cudaEvent_t b = ctx->takeEvent(); cudaEventRecord(b, ctx->stream); ctx->profRecs.push_back({name, a, b}); } }
};
void profileReport(char* buf, int bufLen, bool reset);

Context* acquireContext();            // thread-safe; creates stream/arena on first use; nullptr if no CUDA device
void releaseContext(Context* ctx);
bool cudaOk(cudaError_t e, const char* what);   // logs and returns false on error

enum PtrKind { PTR_HOST_PAGEABLE, PTR_HOST_PINNED, PTR_DEVICE };
PtrKind classifyPointer(const void* p);

// ---------------------------------------------------------------------------------------------
// band-level device work.  All pointers named d* are device pointers valid on ctx->stream.

struct BandMaskState {      // validity of the band being coded and of the previous band (Lerc.cpp:659-741)
  uint8_t* dBits = nullptr;       // MSB-first bit mask, (nPix+7)/8 bytes (BitMask.h:48-67); valid even if all pixels are valid
  uint8_t* dPrevBits = nullptr;
  int numValid = 0;
  bool havePrev = false;
  int pendingFill = -1;           // decode: dBits is to be filled with this byte (all valid / all invalid) before its next use, -1 = dBits is current
};

struct EncodeBandArgs {
  int dt, nDepth, nCols, nRows;
  const void* dData;              // this band's pixels on the device
  const uint8_t* dValidBytes;     // this band's byte mask on the device, or nullptr
  double maxZErr;
  int iBand, nBands, nMasks;
  int version = 6;                // codec version to write: 6, or 2..5 (Lerc::EncodeInternal_v5, Lerc.cpp:526-624)
  // a band that went through prefilterNoData (caller-supplied noData value, Lerc.cpp:687-711): NaN / noData are already moved to the
  // mask or remapped, maxZErr is final, and the all-integer verdict and the header's noData fields are given
  bool prefiltered = false, isAllInt = false, passNoData = false;
  double noDataVal = 0, noDataOrig = 0;
  bool anyMaskModified;           // in/out across bands (Lerc.cpp:714-720)
  const void* hData = nullptr;    // in: the band's pixels in HOST memory when they have not been copied to dData yet (the single-pass encoder
                                  //     copies strip by strip while it codes; everybody else calls stageBand first)
  uint8_t* hOut = nullptr;        // in: host destination of this band's blob (one-band calls with a host output buffer), else nullptr
  bool hostCopied = false;        // out: the band's blob is in hOut already
  uint8_t* dOut;                  // device output buffer (whole multi-band blob), may be nullptr for size-only
  uint8_t* fillEnd = nullptr;     // in: end of the caller's device buffer when the encoder may zero-fill behind the blob itself (one band)
  bool tailFilled = false;        // out: it did
  size_t outCapacity;             // bytes available in dOut from outOffset
  size_t outOffset;               // where this band's blob starts
};

// Encodes (or only sizes, if a.dOut == nullptr) one band.  Returns Ok and the band blob size.
ErrCode encodeBand(Context* ctx, EncodeBandArgs& a, BandMaskState& ms, uint32_t& bandBytes);

struct DecodeBandArgs {
  int dt, nDepth, nCols, nRows;
  const uint8_t* dBlob;           // device pointer to this band's blob
  size_t avail;                   // bytes available from dBlob
  HeaderInfo hd;                  // already parsed on the host
  const uint8_t* hBlob;           // the same band blob in host memory when the caller's blob is a host buffer, else nullptr
  const ByteSource* src = nullptr; size_t srcOff = 0;   // the caller's view of the whole blob and this band's offset in it
  void* dData;                    // device output for this band
  uint8_t* dValidBytes;           // device byte mask output for this band or nullptr
  // host-resident calls: bytes of hBlob that are NOT yet in dBlob (decodeBand copies them itself; the stream decoder strip by strip,
  // overlapped with the decoding) and the caller's host array for the band (the stream decoder sends finished rows there while
  // it still decodes; hostCopied says it did)
  size_t pendingBlob = 0;
  uint8_t* hOut = nullptr;
  bool hostCopied = false;
};
ErrCode decodeBand(Context* ctx, DecodeBandArgs& a, BandMaskState& ms);

// Per-device "done once" flags and cached integers (function attributes and occupancy are per device; contexts are pooled per device)
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool need(int dev) const { return !((mask.load(std::memory_order_relaxed) >> (dev & 63)) & 1ull); }
  void done(int dev) { mask.fetch_or(1ull << (dev & 63), std::memory_order_relaxed); }
};
struct DeviceInt {
  std::atomic<int> v[64];
  int get(int dev) const { return v[dev & 63].load(std::memory_order_relaxed); }
  void set(int dev, int x) { v[dev & 63].store(x, std::memory_order_relaxed); }
};
inline int smCountOf(int dev) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n > 0 ? n : 1; }

// Strip schedule of a pipelined host-resident call: `bytes` of payload in `nUnits` units (block rows / stream chunks).  Small strips
// at both ends (the first kernel starts early, the last copy back is short), strips up to 8x larger in between (fewer launches
// and synchronisations).  bounds[0..n] are unit indices, returns n <= maxStrips (1 = not worth pipelining).
constexpr int kMaxStrips = 16;
inline int stripSchedule(size_t bytes, int nUnits, int bounds[kMaxStrips + 1]) {
  static const int smallLog2 = [] { const char* e = std::getenv("LERC_B200_STRIP_LOG2"); const int v = e ? std::atoi(e) : 0; return (v >= 10 && v <= 30) ? v : 20; }();   // smallest strip ~1 MB
  bounds[0] = 0; bounds[1] = nUnits;
  auto weight = [](int i, int n_) { return 1ll << std::min(std::min(i, n_ - 1 - i), 3); };
  auto weights = [&](int n_) { long long t = 0; for (int i = 0; i < n_; i++) t += weight(i, n_); return t; };
  int n = 1;
  for (int cand = 2; cand <= std::min(kMaxStrips, nUnits); cand++) {   // as many strips as keep the smallest one above the floor
    if ((long long)(bytes >> smallLog2) < weights(cand)) break;
    n = cand;
  }
  if (n < 2) return 1;
  const long long total = weights(n);
  long long acc = 0;
  for (int i = 0; i < n; i++) { acc += weight(i, n); bounds[i + 1] = (int)((long long)nUnits * acc / total); }
  bounds[n] = nUnits;
  int m = 0;                                                           // drop empty strips
  for (int i = 1; i <= n; i++) if (bounds[i] > bounds[m]) bounds[++m] = bounds[i];
  return std::max(m, 1);
}

// Tile batch (include/lerc_b200.h, lerc_tiles_encode.cuh / lerc_tiles_decode.cuh): every tileRows x tileCols window of a
// one-band raster is its own Lerc2 blob.  hOffsets: host array of nTiles + 1 byte offsets into dOut / dBlobs.
struct TilesGeom {
  int dt, nCols, nRows, tileCols, tileRows, nImgX, nImgY;
  long long nImg() const { return (long long)nImgX * nImgY; }
  int rowsOf(long long img) const { const int iy = (int)(img / nImgX); const int r = nRows - iy * tileRows; return r < tileRows ? r : tileRows; }
  int colsOf(long long img) const { const int ix = (int)(img % nImgX); const int c = nCols - ix * tileCols; return c < tileCols ? c : tileCols; }
};
ErrCode encodeTiles(Context* ctx, int dt, int nCols, int nRows, int tileCols, int tileRows, const void* dData, double maxZErr,
                    uint8_t* dOut, size_t outCap, unsigned long long* hOffsets);
ErrCode decodeTiles(Context* ctx, int dt, int nCols, int nRows, int tileCols, int tileRows, const uint8_t* dBlobs, size_t blobBytes,
                    const unsigned long long* hOffsets, void* dData);

// Lerc::FilterNoDataAndNaN / Lerc::FilterNoData (Lerc.cpp:1378-1552, :1241-1374) for a band with a caller-supplied noData value.
// dData / dMaskBytes are private device copies of the band and its byte mask and are modified in place.
struct NoDataVerdict { double maxZErr, noDataVal; bool maskModified, needNoData, isAllInt; };
ErrCode prefilterNoData(Context* ctx, int dt, void* dData, uint8_t* dMaskBytes, long long nPix, int nDepth, double maxZErr, double noDataOrig,
                        NoDataVerdict& out);
// decode side, Lerc::RemapNoData (Lerc.cpp:1046-1076): valid pixels' values equal to (T)from become (T)to
void launchRemapNoData(Context* ctx, int dt, void* dData, const uint8_t* dBits, long long nPix, int nDepth, double from, double to);

// misc device utilities (lerc_mask.cu)
void launchConvertToDouble(Context* ctx, const void* dSrc, int dt, size_t n, double* dDst);   // in-place safe back-to-front

// library statistics for bench.py / tests (lerc_b200.h)
struct Stats { uint64_t kernelLaunches, encodeCalls, decodeCalls, fastPathEncodes, fastPathDecodes; };
Stats& globalStats();

}  // namespace lerc
