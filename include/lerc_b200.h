/*
 * lerc_b200.h -- entry points lerc_b200 adds next to the reference C API (Lerc_c_api.h).
 * Plain C ABI: pointers and sizes only, no CUDA or torch types in the signatures.
 *
 * Nothing here replaces a reference interface; these are the additive extensions announced in
 * (i) the 12 reference functions accept CUDA device pointers for every
 * buffer argument (no new symbol needed), (ii) the caller may supply the CUDA stream the kernels run
 * on, so device-resident pipelines can be timed and ordered with CUDA events, (iii) counters that
 * let a harness prove the CUDA path ran (number of kernel launches, calls).
 */
#ifndef LERC_B200_H
#define LERC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
  #define LERC_B200_API __attribute__((visibility("default")))
#else
  #define LERC_B200_API
#endif

/* Device pointers and ordering: without lerc_b200_set_stream the library works on a private NON-BLOCKING stream, so
 * device buffers handed to lerc_* must be complete (synchronise the producing stream first); results are complete
 * when the call returns.  With lerc_b200_set_stream the calls are ordered on the given stream like any other work.
 * Device buffers are read in aligned 16-byte chunks: a chunk holding the first or last bytes of a buffer may include up to 15
 * bytes outside it (never across a page, never used); pad device blobs to a multiple of 16 bytes if a memory checker objects.
 * Host buffers: pass pinned memory to get the copy / compute overlap inside the call (strips; LERC_B200_STRIP_LOG2).
 *
 * Use `cudaStream` (a cudaStream_t passed as void*) for all work issued by the calling thread's next
 * lerc_* calls when enable != 0; enable == 0 returns to the library's private non-blocking stream.
 * Thread-local.  The lerc_* calls still return only after their results are complete. */
LERC_B200_API void lerc_b200_set_stream(void* cudaStream, int enable);

/* out[0] = kernels launched so far, out[1] = encode calls, out[2] = decode calls,
 * out[3] = encodes that took the fused all-valid fast path, out[4] = decodes that did. */
LERC_B200_API void lerc_b200_get_stats(unsigned long long* out, int n);

/* Per-kernel timing: while enabled, every kernel launch of this library is bracketed by a CUDA event pair on
 * its stream.  lerc_b200_get_profile() writes lines "name<TAB>launches<TAB>total_ms" into buf. */
LERC_B200_API void lerc_b200_profile(int enable);
LERC_B200_API void lerc_b200_get_profile(char* buf, int bufLen, int reset);

/* ---- Tile batch (SURVEY.md 8(b) "batched tile entry point", BASELINE config 5) ------------------------------------------
 * The reference codes one image per lerc_encode / lerc_decode call (Lerc_c_api.h:141-154, :238-252); callers that store
 * rasters as tiles (MRF, GeoTIFF/LERC) loop over the tiles.  These entry points do that loop on the GPU in one pass:
 * the raster pData[nRows][nCols] (one band, nDepth 1, every pixel valid, row-major, no padding) is cut into windows of
 * tileRows x tileCols pixels (edge windows are smaller), tile t = ty * ceil(nCols / tileCols) + tx.  Every tile becomes
 * (encode) / is read from (decode) its own standard Lerc2 blob -- byte for byte the blob lerc_encode writes for that
 * window alone -- stored back to back: blob t occupies bytes [pTileOffsets[t], pTileOffsets[t + 1]) of the buffer
 * (nTiles + 1 offsets; pTileOffsets[0] == 0).  All pointers may be host or CUDA device pointers.
 * Status codes are those of Lerc_types.h.  BufferTooSmall when outBufferSize cannot hold the blobs;
 * lerc_b200_tilesMaxBytes() is a size that always suffices. */
LERC_B200_API unsigned long long lerc_b200_tilesMaxBytes(unsigned int dataType, int nCols, int nRows, int tileCols, int tileRows);
LERC_B200_API unsigned int lerc_b200_encodeTiles(const void* pData, unsigned int dataType, int nCols, int nRows, int tileCols, int tileRows,
                                                 double maxZErr, unsigned char* pOutBuffer, unsigned long long outBufferSize,
                                                 unsigned long long* pTileOffsets, unsigned long long* nBytesWritten);
LERC_B200_API unsigned int lerc_b200_decodeTiles(const unsigned char* pBlobs, unsigned long long blobBytes, const unsigned long long* pTileOffsets,
                                                 unsigned int dataType, int nCols, int nRows, int tileCols, int tileRows, void* pData);

LERC_B200_API const char* lerc_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
