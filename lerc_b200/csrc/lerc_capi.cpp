// lerc_capi.cpp -- the C ABI of libLerc.so.4 (include/Lerc_c_api.h) and the lerc_b200 extensions
// (include/lerc_b200.h).  Replaces the reference's Lerc_c_api_impl.cpp:33-305 + the band loops of
// Lerc::EncodeInternal (Lerc.cpp:628-789) and Lerc::DecodeTempl (Lerc.cpp:397-521): argument checks,
// host/device pointer handling, band iteration.  All pixel work happens in the CUDA translation units.
#include "../../include/Lerc_c_api.h"
#include "../../include/lerc_b200.h"
#include "lerc_internal.h"
#include <cstring>
#include <climits>
#include <algorithm>
#include <vector>

using namespace lerc;

namespace {

thread_local cudaStream_t tlsUserStream = nullptr;
thread_local bool tlsUseUserStream = false;

struct ContextGuard {
  Context* ctx;
  cudaStream_t saved = nullptr;
  ContextGuard() : ctx(acquireContext()) {
    if (ctx && tlsUseUserStream) { saved = ctx->stream; ctx->stream = tlsUserStream; }
  }
  ~ContextGuard() {
    if (ctx) { if (saved) { if (ctx->drainOnRelease) cudaStreamSynchronize(ctx->stream); ctx->stream = saved; ctx->drainOnRelease = false; } releaseContext(ctx); }   // (the caller's stream is drained like the context's own)
  }
};

bool dimsOk(int nDepth, int nCols, int nRows, size_t elem) {            // Lerc.cpp:1622-1639
  if (nDepth <= 0 || nCols <= 0 || nRows <= 0) return false;
  const uint64_t nPix = (uint64_t)nRows * (uint64_t)nCols, lim = (uint64_t)INT_MAX;
  return !(nPix > lim || elem * (uint64_t)nDepth > lim || elem * (uint64_t)nDepth * nPix > lim);
}

bool anyNoData(const unsigned char* pUsesNoData, int nBands) {
  if (!pUsesNoData) return false;
  for (int i = 0; i < nBands; i++) if (pUsesNoData[i]) return true;
  return false;
}

// Brings `bytes` at host/device pointer `p` onto the device (no copy if it already is there).
const void* toDevice(Context* ctx, const void* p, size_t bytes, PtrKind kind, void* dScratch) {
  if (kind == PTR_DEVICE) return p;
  if (!cudaOk(cudaMemcpyAsync(dScratch, p, bytes, cudaMemcpyHostToDevice, ctx->stream), "H2D")) return nullptr;
  return dScratch;
}

lerc_status encodeImpl(const void* pData, int version, unsigned dataType, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                       const unsigned char* pValidBytes, double maxZErr, unsigned char* pOut, unsigned outSize,
                       unsigned* nWritten, unsigned* nNeeded, bool sizeOnly, const unsigned char* pUsesNoData, const double* noDataValues = nullptr) {
  if (!pData || dataType >= (unsigned)DT_Undefined || nDepth <= 0 || nCols <= 0 || nRows <= 0 || nBands <= 0 || maxZErr < 0) return WrongParam;
  if (!sizeOnly && (!pOut || !outSize)) return WrongParam;
  if (!(nMasks == 0 || nMasks == 1 || nMasks == nBands) || (nMasks > 0 && !pValidBytes)) return WrongParam;
  if (version > 6 || (version >= 0 && version < 2)) return WrongParam;   // Lerc2::SetEncoderToOldVersion (Lerc2.cpp:52-63); any negative value = current
  if (version < 0) version = 6;
  const bool haveNoData = anyNoData(pUsesNoData, nBands);
  if (haveNoData && (!noDataValues || version <= 5)) return WrongParam;   // Lerc.cpp:649-652, :378-383 (no noData before codec version 6)
  const size_t ts = (size_t)typeSize((int)dataType);
  if (!dimsOk(nDepth, nCols, nRows, ts)) return DimensionsTooLarge;
  if (version < 4 && nDepth > 1) return Failed;                          // Lerc2::Set refuses nDepth > 1 before codec version 4 (Lerc2.cpp:85-86)

  ContextGuard g;
  Context* ctx = g.ctx;
  if (!ctx) return Failed;
  globalStats().encodeCalls++;

  const size_t nPix = (size_t)nCols * (size_t)nRows, nElemBytes = nPix * (size_t)nDepth * ts, nBits = (nPix + 7) >> 3;
  const PtrKind kData = classifyPointer(pData), kMask = pValidBytes ? classifyPointer(pValidBytes) : PTR_DEVICE;
  const PtrKind kOut = (!sizeOnly) ? classifyPointer(pOut) : PTR_DEVICE;

  void* dBandScratch = kData != PTR_DEVICE ? ctx->arena.alloc(nElemBytes) : nullptr;
  void* dMaskScratch = (pValidBytes && kMask != PTR_DEVICE) ? ctx->arena.alloc(nPix) : nullptr;
  if ((kData != PTR_DEVICE && !dBandScratch) || (pValidBytes && kMask != PTR_DEVICE && !dMaskScratch)) return Failed;

  // device output: the caller's buffer if it is device memory, else a staging buffer bounded by the worst case
  uint8_t* dOut = nullptr; size_t dOutCap = 0;
  if (!sizeOnly) {
    if (kOut == PTR_DEVICE) { dOut = pOut; dOutCap = outSize; }
    else {
      const size_t perBand = 90 + 4 + nBits + nBits / 8000 + 64 + 2 * (size_t)nDepth * ts + 2 + nElemBytes + 16;
      dOutCap = std::min<size_t>((size_t)outSize, perBand * (size_t)nBands);
      dOut = (uint8_t*)ctx->arena.alloc(dOutCap + 16);
      if (!dOut) return Failed;
    }
  }
  BandMaskState ms;
  ms.dPrevBits = (uint8_t*)ctx->arena.alloc(nBits);
  if (!ms.dPrevBits) return Failed;

  size_t offset = 0;
  uint64_t need = 0;
  bool anyModified = false, tailFilled = false, hostCopied = false;
  for (int b = 0; b < nBands; b++) {
    EncodeBandArgs a;
    const size_t arenaMark = ctx->arena.used, pinnedMark = ctx->pinnedUsed;
    a.dt = (int)dataType; a.nDepth = nDepth; a.nCols = nCols; a.nRows = nRows;
    // a host band is not copied here: encodeBand copies it itself (the single-pass encoder strip by strip, overlapped with the coding)
    if (kData == PTR_DEVICE) a.dData = (const uint8_t*)pData + nElemBytes * (size_t)b;
    else { a.dData = dBandScratch; a.hData = (const uint8_t*)pData + nElemBytes * (size_t)b; }
    a.hOut = (!sizeOnly && kOut != PTR_DEVICE && nBands == 1) ? pOut : nullptr;
    a.dValidBytes = nullptr;
    if (nMasks > 0) a.dValidBytes = (const uint8_t*)toDevice(ctx, pValidBytes + (nMasks > 1 ? nPix * (size_t)b : 0), nPix, kMask, dMaskScratch);
    if (!a.dData || (nMasks > 0 && !a.dValidBytes)) return Failed;
    a.maxZErr = maxZErr; a.iBand = b; a.nBands = nBands; a.nMasks = nMasks; a.anyMaskModified = anyModified; a.version = version;
    if (haveNoData && pUsesNoData[b]) {
      // a caller-supplied noData value for this band (Lerc.cpp:687-711): the filter works on private copies of band and mask
      void* dCopy = ctx->arena.alloc(nElemBytes);
      uint8_t* dMaskCopy = (uint8_t*)ctx->arena.alloc(nPix);
      if (!dCopy || !dMaskCopy) return Failed;
      if (!cudaOk(a.hData ? cudaMemcpyAsync(dCopy, a.hData, nElemBytes, cudaMemcpyHostToDevice, ctx->stream)
                          : cudaMemcpyAsync(dCopy, a.dData, nElemBytes, cudaMemcpyDeviceToDevice, ctx->stream), "band copy")) return Failed;
      a.hData = nullptr;
      if (a.dValidBytes) { if (!cudaOk(cudaMemcpyAsync(dMaskCopy, a.dValidBytes, nPix, cudaMemcpyDeviceToDevice, ctx->stream), "mask copy")) return Failed; }
      else cudaMemsetAsync(dMaskCopy, 1, nPix, ctx->stream);
      NoDataVerdict nd;
      const ErrCode pe = prefilterNoData(ctx, (int)dataType, dCopy, dMaskCopy, (long long)nPix, nDepth, maxZErr, noDataValues[b], nd);
      if (pe != Ok) return pe;
      a.dData = dCopy; a.dValidBytes = dMaskCopy; a.maxZErr = nd.maxZErr;
      a.prefiltered = true; a.isAllInt = nd.isAllInt; a.passNoData = nd.needNoData; a.noDataVal = nd.noDataVal; a.noDataOrig = noDataValues[b];
      if (nd.maskModified) a.anyMaskModified = true;
    }
    a.dOut = sizeOnly ? nullptr : dOut; a.outOffset = offset; a.outCapacity = sizeOnly ? 0 : (dOutCap > offset ? dOutCap - offset : 0);
    a.fillEnd = (!sizeOnly && kOut == PTR_DEVICE && nBands == 1) ? pOut + outSize : nullptr;   // the single-pass encoder zero-fills behind the blob itself
    uint32_t bandBytes = 0;
    const ErrCode e = encodeBand(ctx, a, ms, bandBytes);
    if (e == BufferTooSmall && !sizeOnly && kOut != PTR_DEVICE && dOutCap < (size_t)outSize) return Failed;   // our bound was wrong: never expected
    if (e != Ok) return e;
    anyModified = a.anyMaskModified;
    tailFilled = a.tailFilled;
    hostCopied = a.hostCopied;
    if (need + bandBytes > (uint64_t)UINT_MAX) return DimensionsTooLarge;   // Lerc.cpp:757-758
    need += bandBytes;
    offset += bandBytes;
    // per-band scratch is recycled; everything allocated before the band loop (and the mask of the previous band) stays.
    // (After the last band of a writing call the synchronisation below, behind the tail fill / blob copy, covers it.)
    if (b + 1 < nBands || sizeOnly) {
      if (!cudaOk(cudaStreamSynchronize(ctx->stream), "band sync")) return Failed;
      if (ctx->arena.retired.empty()) ctx->arena.used = arenaMark;
      ctx->pinnedUsed = pinnedMark;
    }
  }
  if (nNeeded) *nNeeded = (unsigned)need;
  if (sizeOnly) return Ok;

  // the API zero-fills the whole output buffer before writing (Lerc.cpp:374): blob, then zeros
  if (kOut == PTR_DEVICE) {
    if (!tailFilled && outSize > offset && !cudaOk(cudaMemsetAsync(pOut + offset, 0, outSize - offset, ctx->stream), "memset tail")) return Failed;
    // a device blob on the caller's stream (lerc_b200_set_stream) is complete in stream order (lerc_b200.h); otherwise on return
    if (!tlsUseUserStream && !cudaOk(cudaStreamSynchronize(ctx->stream), "sync")) return Failed;
    ctx->drainOnRelease = false;
  } else {
    if (!hostCopied && !cudaOk(cudaMemcpyAsync(pOut, dOut, offset, cudaMemcpyDeviceToHost, ctx->stream), "D2H blob")) return Failed;
    if (outSize > offset) std::memset(pOut + offset, 0, outSize - offset);
    if (!cudaOk(cudaStreamSynchronize(ctx->stream), "sync")) return Failed;
    ctx->drainOnRelease = false;
  }
  *nWritten = (unsigned)offset;
  return Ok;
}

lerc_status decodeImpl(const unsigned char* pBlob, unsigned blobSize, int nMasks, unsigned char* pValidBytes, int nDepth, int nCols,
                       int nRows, int nBands, unsigned dataType, void* pData, bool toDouble, unsigned char* pUsesNoData, double* noDataValues) {
  if (!pBlob || !blobSize || !pData || nDepth <= 0 || nCols <= 0 || nRows <= 0 || nBands <= 0) return WrongParam;
  if (!toDouble && dataType >= (unsigned)DT_Undefined) return WrongParam;
  if (!(nMasks == 0 || nMasks == 1 || nMasks == nBands) || (nMasks > 0 && !pValidBytes)) return WrongParam;

  const PtrKind kBlob = classifyPointer(pBlob);
  ContextGuard g;                                                       // (before the header parse: device blobs are read on the call's stream)
  Context* ctx = g.ctx;
  ByteSource src; src.base = pBlob; src.size = blobSize; src.onDevice = kBlob == PTR_DEVICE;
  if (src.onDevice) { if (!ctx) return Failed; src.stream = ctx->stream; }
  BlobInfo li;
  ErrCode e = getBlobInfo(src, li, nullptr, nullptr, 0);                // fast; does most checks (Lerc.cpp:418)
  if (e != Ok) return e;
  if (toDouble) {                                                       // Lerc_c_api_impl.cpp:268-277
    if (li.nDepth != nDepth || li.nCols != nCols || li.nRows != nRows || li.nBands != nBands) return Failed;
    dataType = (unsigned)li.dt;
  }
  const size_t ts = (size_t)typeSize((int)dataType);
  if (!dimsOk(nDepth, nCols, nRows, ts)) return DimensionsTooLarge;
  if (nMasks < li.nMasks || nBands > li.nBands) return WrongParam;     // Lerc.cpp:423-428
  const bool wantNoData = li.nUsesNoDataValue && nDepth > 1;
  if (wantNoData) {                                                    // Lerc.cpp:431-445
    if (!pUsesNoData || !noDataValues) return HasNoData;
    std::memset(pUsesNoData, 0, (size_t)nBands);
    std::memset(noDataValues, 0, (size_t)nBands * sizeof(double));
  }

  if (!ctx) return Failed;
  globalStats().decodeCalls++;

  const size_t nPix = (size_t)nCols * (size_t)nRows, nElem = nPix * (size_t)nDepth, nBits = (nPix + 7) >> 3;
  const PtrKind kData = classifyPointer(pData), kMask = pValidBytes ? classifyPointer(pValidBytes) : PTR_DEVICE;

  const uint8_t* dBlob = pBlob;
  if (kBlob != PTR_DEVICE) {
    uint8_t* d = (uint8_t*)ctx->arena.alloc((size_t)blobSize + 64);
    // (a single-band blob is left to decodeBand: its stream decoder copies and decodes it strip by strip)
    if (!d || (nBands > 1 && !cudaOk(cudaMemcpyAsync(d, pBlob, blobSize, cudaMemcpyHostToDevice, ctx->stream), "H2D blob"))) return Failed;
    dBlob = d;
  }
  const bool blobPending = kBlob != PTR_DEVICE && nBands == 1;
  BandMaskState ms;
  ms.dBits = (uint8_t*)ctx->arena.alloc(nBits);
  if (!ms.dBits) return Failed;
  // typed band buffer: the caller's memory when it is device memory of the right type, else scratch
  const bool direct = kData == PTR_DEVICE && !(toDouble && dataType != (unsigned)DT_Double);
  void* dBandScratch = direct ? nullptr : ctx->arena.alloc(nElem * ts);
  double* dDoubleScratch = (toDouble && dataType != (unsigned)DT_Double && kData != PTR_DEVICE) ? (double*)ctx->arena.alloc(nElem * 8) : nullptr;
  uint8_t* dMaskScratch = (pValidBytes && kMask != PTR_DEVICE) ? (uint8_t*)ctx->arena.alloc(nPix) : nullptr;
  if ((!direct && !dBandScratch) || (pValidBytes && kMask != PTR_DEVICE && !dMaskScratch)) return Failed;

  size_t pos = 0;
  for (int b = 0; b < nBands; b++) {
    HeaderInfo hd;
    uint8_t head[96];
    const size_t avail = blobSize > pos ? blobSize - pos : 0, take = std::min(avail, sizeof head);
    if (pos >= blobSize || take < 14 || !src.fetch(pos, take, head) || !readHeader(head, take, hd)) break;   // Lerc.cpp:453
    if (hd.nDepth != nDepth || hd.nCols != nCols || hd.nRows != nRows || pos + (size_t)hd.blobSize > blobSize) return Failed;

    DecodeBandArgs a;
    a.dt = (int)dataType; a.nDepth = nDepth; a.nCols = nCols; a.nRows = nRows;
    a.dBlob = dBlob + pos; a.avail = blobSize - pos; a.hd = hd; a.hBlob = kBlob != PTR_DEVICE ? pBlob + pos : nullptr;
    a.src = &src; a.srcOff = pos;
    if (blobPending) a.pendingBlob = (size_t)hd.blobSize;
    if (!direct && !(toDouble && dataType != (unsigned)DT_Double)) a.hOut = (uint8_t*)pData + nElem * ts * (size_t)b;
    a.dData = direct ? (void*)((uint8_t*)pData + nElem * ts * (size_t)b) : dBandScratch;
    const bool wantMask = b < nMasks;
    a.dValidBytes = wantMask ? (kMask == PTR_DEVICE ? pValidBytes + nPix * (size_t)b : dMaskScratch) : nullptr;
    const size_t arenaMark = ctx->arena.used, pinnedMark = ctx->pinnedUsed;
    e = decodeBand(ctx, a, ms);
    if (e != Ok) return e;
    if (wantNoData) {                                                  // Lerc.cpp:472-479
      pUsesNoData[b] = hd.bPassNoDataValues ? 1 : 0;
      noDataValues[b] = hd.noDataValOrig;
      if (hd.bPassNoDataValues && hd.noDataVal != hd.noDataValOrig)
        launchRemapNoData(ctx, (int)dataType, a.dData, ms.dBits, (long long)nPix, nDepth, hd.noDataVal, hd.noDataValOrig);
    }

    if (!direct) {
      if (toDouble && dataType != (unsigned)DT_Double) {
        double* dst = kData == PTR_DEVICE ? (double*)pData + nElem * (size_t)b : dDoubleScratch;
        launchConvertToDouble(ctx, dBandScratch, (int)dataType, nElem, dst);
        if (kData != PTR_DEVICE && !cudaOk(cudaMemcpyAsync((double*)pData + nElem * (size_t)b, dst, nElem * 8, cudaMemcpyDeviceToHost, ctx->stream), "D2H")) return Failed;
      } else if (!a.hostCopied && !cudaOk(cudaMemcpyAsync((uint8_t*)pData + nElem * ts * (size_t)b, dBandScratch, nElem * ts, cudaMemcpyDeviceToHost, ctx->stream), "D2H")) return Failed;
    }
    if (wantMask && kMask != PTR_DEVICE &&
        !cudaOk(cudaMemcpyAsync(pValidBytes + nPix * (size_t)b, dMaskScratch, nPix, cudaMemcpyDeviceToHost, ctx->stream), "D2H mask")) return Failed;
    if (!cudaOk(cudaStreamSynchronize(ctx->stream), "band sync")) return Failed;
    if (ctx->arena.retired.empty()) ctx->arena.used = arenaMark;
    ctx->pinnedUsed = pinnedMark;
    pos += (size_t)hd.blobSize;
  }
  ctx->drainOnRelease = false;                                         // (every band ended with a synchronisation)
  return Ok;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

lerc_status lerc_computeCompressedSizeForVersion(const void* pData, int codecVersion, unsigned int dataType, int nDepth, int nCols, int nRows,
                                                 int nBands, int nMasks, const unsigned char* pValidBytes, double maxZErr, unsigned int* numBytes) {
  if (!numBytes) return WrongParam;
  *numBytes = 0;
  unsigned w = 0;
  return encodeImpl(pData, codecVersion, dataType, nDepth, nCols, nRows, nBands, nMasks, pValidBytes, maxZErr, nullptr, 0, &w, numBytes, true, nullptr);
}

lerc_status lerc_computeCompressedSize(const void* pData, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                                       const unsigned char* pValidBytes, double maxZErr, unsigned int* numBytes) {
  return lerc_computeCompressedSizeForVersion(pData, -1, dataType, nDepth, nCols, nRows, nBands, nMasks, pValidBytes, maxZErr, numBytes);
}

lerc_status lerc_encodeForVersion(const void* pData, int codecVersion, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands,
                                  int nMasks, const unsigned char* pValidBytes, double maxZErr, unsigned char* pOutBuffer,
                                  unsigned int outBufferSize, unsigned int* nBytesWritten) {
  if (!nBytesWritten) return WrongParam;
  *nBytesWritten = 0;
  unsigned need = 0;
  return encodeImpl(pData, codecVersion, dataType, nDepth, nCols, nRows, nBands, nMasks, pValidBytes, maxZErr, pOutBuffer, outBufferSize,
                    nBytesWritten, &need, false, nullptr);
}

lerc_status lerc_encode(const void* pData, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                        const unsigned char* pValidBytes, double maxZErr, unsigned char* pOutBuffer, unsigned int outBufferSize,
                        unsigned int* nBytesWritten) {
  return lerc_encodeForVersion(pData, -1, dataType, nDepth, nCols, nRows, nBands, nMasks, pValidBytes, maxZErr, pOutBuffer, outBufferSize, nBytesWritten);
}

lerc_status lerc_computeCompressedSize_4D(const void* pData, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                                          const unsigned char* pValidBytes, double maxZErr, unsigned int* numBytes,
                                          const unsigned char* pUsesNoData, const double* noDataValues) {
  if (!numBytes) return WrongParam;
  *numBytes = 0;
  unsigned w = 0;
  return encodeImpl(pData, -1, dataType, nDepth, nCols, nRows, nBands, nMasks, pValidBytes, maxZErr, nullptr, 0, &w, numBytes, true, pUsesNoData, noDataValues);
}

lerc_status lerc_encode_4D(const void* pData, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                           const unsigned char* pValidBytes, double maxZErr, unsigned char* pOutBuffer, unsigned int outBufferSize,
                           unsigned int* nBytesWritten, const unsigned char* pUsesNoData, const double* noDataValues) {
  if (!nBytesWritten) return WrongParam;
  *nBytesWritten = 0;
  unsigned need = 0;
  return encodeImpl(pData, -1, dataType, nDepth, nCols, nRows, nBands, nMasks, pValidBytes, maxZErr, pOutBuffer, outBufferSize,
                    nBytesWritten, &need, false, pUsesNoData, noDataValues);
}

lerc_status lerc_getBlobInfo(const unsigned char* pLercBlob, unsigned int blobSize, unsigned int* infoArray, double* dataRangeArray,
                             int infoArraySize, int dataRangeArraySize) {
  if (!pLercBlob || !blobSize || (!infoArray && !dataRangeArray) || (infoArraySize <= 0 && dataRangeArraySize <= 0)) return WrongParam;
  ByteSource src; src.base = pLercBlob; src.size = blobSize; src.onDevice = classifyPointer(pLercBlob) == PTR_DEVICE;
  BlobInfo li;
  const ErrCode e = getBlobInfo(src, li, nullptr, nullptr, 0);
  if (e != Ok) return e;
  if (infoArray) {                                                      // Lerc_c_api_impl.cpp:106-133
    const unsigned v[11] = {(unsigned)li.version, (unsigned)li.dt, (unsigned)li.nDepth, (unsigned)li.nCols, (unsigned)li.nRows, (unsigned)li.nBands,
                            (unsigned)li.numValidPixel, li.blobSize, (unsigned)li.nMasks, (unsigned)li.nDepth, (unsigned)li.nUsesNoDataValue};
    for (int i = 0; i < infoArraySize; i++) infoArray[i] = i < 11 ? v[i] : 0;
  }
  if (dataRangeArray) {                                                 // Lerc_c_api_impl.cpp:136-154
    const bool noData = li.nDepth > 1 && li.nUsesNoDataValue > 0;
    const double v[3] = {noData ? -1 : li.zMin, noData ? -1 : li.zMax, li.maxZError};
    for (int i = 0; i < dataRangeArraySize; i++) dataRangeArray[i] = i < 3 ? v[i] : 0;
  }
  return Ok;
}

lerc_status lerc_getDataRanges(const unsigned char* pLercBlob, unsigned int blobSize, int nDepth, int nBands, double* pMins, double* pMaxs) {
  if (!pLercBlob || !blobSize || !pMins || !pMaxs || nDepth <= 0 || nBands <= 0) return WrongParam;
  ByteSource src; src.base = pLercBlob; src.size = blobSize; src.onDevice = classifyPointer(pLercBlob) == PTR_DEVICE;
  BlobInfo li;
  return getBlobInfo(src, li, pMins, pMaxs, (size_t)nDepth * (size_t)nBands);
}

lerc_status lerc_decode_4D(const unsigned char* pLercBlob, unsigned int blobSize, int nMasks, unsigned char* pValidBytes, int nDepth, int nCols,
                           int nRows, int nBands, unsigned int dataType, void* pData, unsigned char* pUsesNoData, double* noDataValues) {
  return decodeImpl(pLercBlob, blobSize, nMasks, pValidBytes, nDepth, nCols, nRows, nBands, dataType, pData, false, pUsesNoData, noDataValues);
}

lerc_status lerc_decode(const unsigned char* pLercBlob, unsigned int blobSize, int nMasks, unsigned char* pValidBytes, int nDepth, int nCols,
                        int nRows, int nBands, unsigned int dataType, void* pData) {
  return lerc_decode_4D(pLercBlob, blobSize, nMasks, pValidBytes, nDepth, nCols, nRows, nBands, dataType, pData, nullptr, nullptr);
}

lerc_status lerc_decodeToDouble_4D(const unsigned char* pLercBlob, unsigned int blobSize, int nMasks, unsigned char* pValidBytes, int nDepth,
                                   int nCols, int nRows, int nBands, double* pData, unsigned char* pUsesNoData, double* noDataValues) {
  return decodeImpl(pLercBlob, blobSize, nMasks, pValidBytes, nDepth, nCols, nRows, nBands, 0, pData, true, pUsesNoData, noDataValues);
}

lerc_status lerc_decodeToDouble(const unsigned char* pLercBlob, unsigned int blobSize, int nMasks, unsigned char* pValidBytes, int nDepth,
                                int nCols, int nRows, int nBands, double* pData) {
  return lerc_decodeToDouble_4D(pLercBlob, blobSize, nMasks, pValidBytes, nDepth, nCols, nRows, nBands, pData, nullptr, nullptr);
}

// ---- lerc_b200 extensions ----------------------------------------------------------------------
void lerc_b200_set_stream(void* cudaStream, int enable) { tlsUserStream = (cudaStream_t)cudaStream; tlsUseUserStream = enable != 0; }

void lerc_b200_get_stats(unsigned long long* out, int n) {
  const Stats& s = globalStats();
  const unsigned long long v[5] = {s.kernelLaunches, s.encodeCalls, s.decodeCalls, s.fastPathEncodes, s.fastPathDecodes};
  for (int i = 0; i < n; i++) out[i] = i < 5 ? v[i] : 0;
}

void lerc_b200_profile(int enable) { gProfileEnabled = enable != 0; }
void lerc_b200_get_profile(char* buf, int bufLen, int reset) { profileReport(buf, bufLen, reset != 0); }

// ---- tile batch (lerc_tiles_encode.cuh / lerc_tiles_decode.cuh) -----------------------------------
unsigned long long lerc_b200_tilesMaxBytes(unsigned int dataType, int nCols, int nRows, int tileCols, int tileRows) {
  if (dataType >= (unsigned)DT_Undefined || nCols <= 0 || nRows <= 0 || tileCols <= 0 || tileRows <= 0) return 0;
  const unsigned long long ts = (unsigned long long)typeSize((int)dataType);
  const unsigned long long nx = ((unsigned long long)nCols + tileCols - 1) / tileCols, ny = ((unsigned long long)nRows + tileRows - 1) / tileRows;
  const unsigned long long bx = ((unsigned long long)std::min(tileCols, nCols) + 7) / 8, by = ((unsigned long long)std::min(tileRows, nRows) + 7) / 8;
  // per tile: header, mask byte count, ranges, two flag bytes, every block raw with its flag byte (the longest stream
  // the single-pass encoder can emit before a smaller coding is chosen), slack for aligned stores
  const unsigned long long perTile = 90 + 4 + 2 * ts + 2 + (unsigned long long)std::min(tileCols, nCols) * std::min(tileRows, nRows) * ts + bx * by + 32;
  return nx * ny * perTile;
}

lerc_status lerc_b200_encodeTiles(const void* pData, unsigned int dataType, int nCols, int nRows, int tileCols, int tileRows, double maxZErr,
                                  unsigned char* pOutBuffer, unsigned long long outBufferSize, unsigned long long* pTileOffsets,
                                  unsigned long long* nBytesWritten) {
  if (nBytesWritten) *nBytesWritten = 0;
  if (!pData || !pOutBuffer || !pTileOffsets || !nBytesWritten || dataType >= (unsigned)DT_Undefined || !(maxZErr >= 0)) return WrongParam;
  if (nCols <= 0 || nRows <= 0 || tileCols <= 0 || tileRows <= 0 || outBufferSize == 0) return WrongParam;
  const size_t ts = (size_t)typeSize((int)dataType);
  if (!dimsOk(1, std::min(tileCols, nCols), std::min(tileRows, nRows), ts)) return DimensionsTooLarge;
  const unsigned long long nImg = (((unsigned long long)nCols + tileCols - 1) / tileCols) * (((unsigned long long)nRows + tileRows - 1) / tileRows);
  if (nImg > 0x7fffffffull) return DimensionsTooLarge;
  ContextGuard g;
  Context* ctx = g.ctx;
  if (!ctx) return Failed;
  globalStats().encodeCalls++;
  const size_t rasterBytes = (size_t)nCols * (size_t)nRows * ts;
  const PtrKind kData = classifyPointer(pData), kOut = classifyPointer(pOutBuffer), kOff = classifyPointer(pTileOffsets);
  const void* dData = pData;
  if (kData != PTR_DEVICE) {
    void* d = ctx->arena.alloc(rasterBytes);
    if (!d || !cudaOk(cudaMemcpyAsync(d, pData, rasterBytes, cudaMemcpyHostToDevice, ctx->stream), "H2D raster")) return Failed;
    dData = d;
  }
  uint8_t* dOut = pOutBuffer; size_t dOutCap = (size_t)outBufferSize;
  if (kOut != PTR_DEVICE) {
    dOutCap = (size_t)std::min<unsigned long long>(outBufferSize, lerc_b200_tilesMaxBytes(dataType, nCols, nRows, tileCols, tileRows));
    dOut = (uint8_t*)ctx->arena.alloc(dOutCap + 16);
    if (!dOut) return Failed;
  }
  std::vector<unsigned long long> hOff((size_t)nImg + 1, 0);
  const ErrCode e = encodeTiles(ctx, (int)dataType, nCols, nRows, tileCols, tileRows, dData, maxZErr, dOut, dOutCap, hOff.data());
  if (e == BufferTooSmall && kOut != PTR_DEVICE && dOutCap < (size_t)outBufferSize) return Failed;   // the bound was wrong: never expected
  if (e != Ok) return e;
  const unsigned long long total = hOff[(size_t)nImg];
  if (kOut != PTR_DEVICE && !cudaOk(cudaMemcpyAsync(pOutBuffer, dOut, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream), "D2H blobs")) return Failed;
  if (kOff == PTR_DEVICE) { if (!cudaOk(cudaMemcpyAsync(pTileOffsets, hOff.data(), hOff.size() * 8, cudaMemcpyHostToDevice, ctx->stream), "H2D offsets")) return Failed; }
  else std::memcpy(pTileOffsets, hOff.data(), hOff.size() * 8);
  if (!cudaOk(cudaStreamSynchronize(ctx->stream), "sync")) return Failed;
  *nBytesWritten = total;
  return Ok;
}

lerc_status lerc_b200_decodeTiles(const unsigned char* pBlobs, unsigned long long blobBytes, const unsigned long long* pTileOffsets,
                                  unsigned int dataType, int nCols, int nRows, int tileCols, int tileRows, void* pData) {
  if (!pBlobs || !blobBytes || !pTileOffsets || !pData || dataType >= (unsigned)DT_Undefined) return WrongParam;
  if (nCols <= 0 || nRows <= 0 || tileCols <= 0 || tileRows <= 0) return WrongParam;
  const size_t ts = (size_t)typeSize((int)dataType);
  if (!dimsOk(1, std::min(tileCols, nCols), std::min(tileRows, nRows), ts)) return DimensionsTooLarge;
  const unsigned long long nImg = (((unsigned long long)nCols + tileCols - 1) / tileCols) * (((unsigned long long)nRows + tileRows - 1) / tileRows);
  if (nImg > 0x7fffffffull) return DimensionsTooLarge;
  ContextGuard g;
  Context* ctx = g.ctx;
  if (!ctx) return Failed;
  globalStats().decodeCalls++;
  const size_t rasterBytes = (size_t)nCols * (size_t)nRows * ts;
  const PtrKind kBlob = classifyPointer(pBlobs), kData = classifyPointer(pData), kOff = classifyPointer(pTileOffsets);
  std::vector<unsigned long long> hOff((size_t)nImg + 1);
  if (kOff == PTR_DEVICE) {
    if (!cudaOk(cudaMemcpyAsync(hOff.data(), pTileOffsets, hOff.size() * 8, cudaMemcpyDeviceToHost, ctx->stream), "D2H offsets") ||
        !cudaOk(cudaStreamSynchronize(ctx->stream), "sync")) return Failed;
  } else std::memcpy(hOff.data(), pTileOffsets, hOff.size() * 8);
  const uint8_t* dBlobs = pBlobs;
  if (kBlob != PTR_DEVICE) {
    uint8_t* d = (uint8_t*)ctx->arena.alloc((size_t)blobBytes + 64);
    if (!d || !cudaOk(cudaMemcpyAsync(d, pBlobs, (size_t)blobBytes, cudaMemcpyHostToDevice, ctx->stream), "H2D blobs")) return Failed;
    dBlobs = d;
  }
  void* dData = pData;
  if (kData != PTR_DEVICE) { dData = ctx->arena.alloc(rasterBytes); if (!dData) return Failed; }
  const ErrCode e = decodeTiles(ctx, (int)dataType, nCols, nRows, tileCols, tileRows, dBlobs, (size_t)blobBytes, hOff.data(), dData);
  if (e != Ok) return e;
  if (kData != PTR_DEVICE && !cudaOk(cudaMemcpyAsync(pData, dData, rasterBytes, cudaMemcpyDeviceToHost, ctx->stream), "D2H raster")) return Failed;
  return cudaOk(cudaStreamSynchronize(ctx->stream), "sync") ? Ok : Failed;
}

const char* lerc_b200_version(void) { return "lerc_b200 0.2 (Lerc2 v2-v6 writer and reader, tile batch; CUDA sm_100a)"; }

}  // extern "C"
