/*
 * Lerc_c_api.h -- the C ABI of libLerc.so.4 as implemented by lerc_b200 (CUDA, sm_100a).
 *
 * This is the drop-in boundary: the 12 entry points below have exactly the names, argument lists
 * and lerc_status semantics of the reference's src/LercLib/include/Lerc_c_api.h:126-380 (Esri/lerc
 * v4.2.0), so existing callers (GDAL, the ctypes wrapper OtherLanguages/Python/lerc/_lerc.py:277-312,
 * P/Invoke, ...) bind to this library without change.  What is behind them is new: the Lerc2
 * micro-block pipeline runs as hand-written CUDA kernels (lerc_b200/csrc).
 *
 * Conventions shared by all functions (reference Lerc_c_api.h:111-124):
 *   - the caller allocates every buffer; nothing is returned that must be freed;
 *   - pixel order is row by row, top-left first, band after band; with nDepth > 1 the nDepth values
 *     of one pixel are adjacent ([RGB, RGB, ...]);
 *   - dataType: 0 char, 1 uchar, 2 short, 3 ushort, 4 int, 5 uint, 6 float, 7 double (Lerc_types.h);
 *   - a validity mask is 1 byte per pixel (1 valid, 0 invalid), nCols*nRows*nMasks bytes,
 *     nMasks in {0, 1, nBands}; a null mask means all pixels are valid;
 *   - return value: 0 Ok, 1 Failed, 2 WrongParam, 3 BufferTooSmall, 4 NaN, 5 HasNoData,
 *     6 DimensionsTooLarge (Lerc_types.h).
 *
 * lerc_b200 additions that do not change the ABI: every data / blob / mask pointer may be a host
 * pointer (pageable or pinned) OR a CUDA device pointer; device-resident buffers are used in place.
 * Further entry points (batched tiles, explicit streams, timing hooks) live in lerc_b200.h.
 *
 * Implemented like the reference: codec versions 2..6 on the write side (lerc_*ForVersion), the
 * maxZErr == 777 bit-plane switch, the noData arguments of the _4D functions, the lossless float
 * (FPL) codec.  Documented deviations (DESIGN.md "Deviations"): Lerc1 blobs are not read (Failed);
 * lerc_decode with a data type different from the blob's returns Failed (the reference
 * reinterprets); the pUsesNoData / noDataValues arrays of the _4D functions are host arrays.
 * tests/test_capi_boundary.py compares every prototype below with the reference's header.
 */
#ifndef LERC_API_INCLUDE_GUARD
#define LERC_API_INCLUDE_GUARD

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_MSC_VER)
  #define LERCDLL_API
#elif defined(LERC_EXPORTS)
  #define LERCDLL_API __attribute__((visibility("default")))
#else
  #define LERCDLL_API
#endif

#define LERC_VERSION_MAJOR 4
#define LERC_VERSION_MINOR 2
#define LERC_VERSION_PATCH 0
#define LERC_AT_LEAST_VERSION(maj, min, patch) \
  (LERC_VERSION_MAJOR > (maj) || (LERC_VERSION_MAJOR == (maj) && (LERC_VERSION_MINOR > (min) || \
  (LERC_VERSION_MINOR == (min) && LERC_VERSION_PATCH >= (patch)))))

  typedef unsigned int lerc_status;

  /* Exact size in bytes lerc_encode() will produce for this input (reference Lerc_c_api.h:126-137). */
  LERCDLL_API lerc_status lerc_computeCompressedSize(
      const void* pData, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands,
      int nMasks, const unsigned char* pValidBytes, double maxZErr, unsigned int* numBytes);

  /* Compress nBands rasters into one blob of concatenated Lerc2 v6 band blobs.  The whole output
   * buffer is zero-filled first; BufferTooSmall if outBufferSize is less than the exact size
   * (reference Lerc_c_api.h:141-154, Lerc.cpp:374, :764). */
  LERCDLL_API lerc_status lerc_encode(
      const void* pData, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands,
      int nMasks, const unsigned char* pValidBytes, double maxZErr,
      unsigned char* pOutBuffer, unsigned int outBufferSize, unsigned int* nBytesWritten);

  /* Same with an explicit codec version: -1 / 6 = current, 2..5 = the older writers (reference :159-187). */
  LERCDLL_API lerc_status lerc_computeCompressedSizeForVersion(
      const void* pData, int codecVersion, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands,
      int nMasks, const unsigned char* pValidBytes, double maxZErr, unsigned int* numBytes);

  LERCDLL_API lerc_status lerc_encodeForVersion(
      const void* pData, int codecVersion, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands,
      int nMasks, const unsigned char* pValidBytes, double maxZErr,
      unsigned char* pOutBuffer, unsigned int outBufferSize, unsigned int* nBytesWritten);

  /* Header-only inspection (host work): infoArray = { version, dataType, nDepth, nCols, nRows, nBands,
   * nValidPixels(band 0), blobSize(total), nMasks, nDepth, nUsesNoDataValue }, dataRangeArray =
   * { zMin, zMax, maxZErrUsed }; both filled up to the sizes given (reference :206-216). */
  LERCDLL_API lerc_status lerc_getBlobInfo(
      const unsigned char* pLercBlob, unsigned int blobSize,
      unsigned int* infoArray, double* dataRangeArray, int infoArraySize, int dataRangeArraySize);

  /* Per band and depth [min, max] without decoding pixels; arrays hold nDepth*nBands doubles (reference :222-232). */
  LERCDLL_API lerc_status lerc_getDataRanges(
      const unsigned char* pLercBlob, unsigned int blobSize, int nDepth, int nBands, double* pMins, double* pMaxs);

  /* Decompress into pData (nDepth*nCols*nRows*nBands values of dataType) and, if pValidBytes is not
   * null, nMasks masks (filled even when all pixels are valid) (reference :238-252). */
  LERCDLL_API lerc_status lerc_decode(
      const unsigned char* pLercBlob, unsigned int blobSize, int nMasks, unsigned char* pValidBytes,
      int nDepth, int nCols, int nRows, int nBands, unsigned int dataType, void* pData);

  /* Decode any pixel type into doubles (reference :260-270). */
  LERCDLL_API lerc_status lerc_decodeToDouble(
      const unsigned char* pLercBlob, unsigned int blobSize, int nMasks, unsigned char* pValidBytes,
      int nDepth, int nCols, int nRows, int nBands, double* pData);

  /* The four "_4D" functions of API v4.0 add per-band noData arguments (reference :305-380).
   * With pUsesNoData == nullptr (or all zeros) they are identical to the functions above. */
  LERCDLL_API lerc_status lerc_computeCompressedSize_4D(
      const void* pData, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands,
      int nMasks, const unsigned char* pValidBytes, double maxZErr, unsigned int* numBytes,
      const unsigned char* pUsesNoData, const double* noDataValues);

  LERCDLL_API lerc_status lerc_encode_4D(
      const void* pData, unsigned int dataType, int nDepth, int nCols, int nRows, int nBands,
      int nMasks, const unsigned char* pValidBytes, double maxZErr,
      unsigned char* pOutBuffer, unsigned int outBufferSize, unsigned int* nBytesWritten,
      const unsigned char* pUsesNoData, const double* noDataValues);

  LERCDLL_API lerc_status lerc_decode_4D(
      const unsigned char* pLercBlob, unsigned int blobSize, int nMasks, unsigned char* pValidBytes,
      int nDepth, int nCols, int nRows, int nBands, unsigned int dataType, void* pData,
      unsigned char* pUsesNoData, double* noDataValues);

  LERCDLL_API lerc_status lerc_decodeToDouble_4D(
      const unsigned char* pLercBlob, unsigned int blobSize, int nMasks, unsigned char* pValidBytes,
      int nDepth, int nCols, int nRows, int nBands, double* pData,
      unsigned char* pUsesNoData, double* noDataValues);

#ifdef __cplusplus
}
#endif
#endif  /* LERC_API_INCLUDE_GUARD */
