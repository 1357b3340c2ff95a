/*
 * lerc_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C (C99) CPU restatement of the Lerc2 codec hot path of Esri/lerc v4.2.0, written from the
 * byte-stream rules in SURVEY.md Appendix A and the behaviour of the reference sources cited
 * function by function below (paths relative to /root/reference).  It is the parity checker for the
 * CUDA product in lerc_b200/: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  The product never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement against
 *   (a) the fixtures shipped in the reference tree (tests/golden/ holds their decoded pixels,
 *       produced by tests/golden/make_golden.py from the unmodified reference), and
 *   (b) the unmodified reference itself compiled to oracle/_ref/libLerc_ref.so (oracle/Makefile),
 *       byte-for-byte on seeded synthetic rasters of all 8 pixel types.
 *
 * Scope: Lerc2 writer and reader for codec versions 2..6 (2..5 = Lerc::EncodeInternal_v5; version 2 with
 * its MSB-first bit stuffing and no checksum); tiling, one-sweep raw, const image, RLE bit mask, per-depth ranges,
 * depth-delta blocks, LUT blocks, 8-bit Huffman / delta-Huffman, Fletcher-32, multi-band concatenation, the _4D calls with
 * per-band noData values (FilterNoDataAndNaN / FilterNoData / RemapNoData), the DECODER of the lossless-float FPL codec.
 * Not restated (SURVEY 8f "next"): the FPL encoder (maxZError == 0 float/double blobs are written as raw/const micro-blocks,
 * which every Lerc2 reader decodes), Lerc1.  (The integer bit-plane mode, maxZErr == 777, is restated.)
 *
 * The exported functions use the reference C API's argument lists (src/LercLib/include/Lerc_c_api.h:126-380)
 * with an `lo_` prefix so one ctypes binding drives all three libraries.
 */
#ifndef LERC_ORACLE_H
#define LERC_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { LO_OK = 0, LO_FAILED = 1, LO_WRONG_PARAM = 2, LO_BUFFER_TOO_SMALL = 3, LO_NAN = 4, LO_HAS_NODATA = 5, LO_DIMS_TOO_LARGE = 6 };
enum { LO_CHAR = 0, LO_BYTE, LO_SHORT, LO_USHORT, LO_INT, LO_UINT, LO_FLOAT, LO_DOUBLE, LO_UNDEFINED };

unsigned lo_computeCompressedSize(const void* data, unsigned dt, int nDepth, int nCols, int nRows, int nBands,
                                  int nMasks, const unsigned char* validBytes, double maxZErr, unsigned* numBytes);
unsigned lo_encode(const void* data, unsigned dt, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                   const unsigned char* validBytes, double maxZErr, unsigned char* out, unsigned outSize, unsigned* nWritten);
unsigned lo_computeCompressedSizeForVersion(const void* data, int version, unsigned dt, int nDepth, int nCols, int nRows,
                                            int nBands, int nMasks, const unsigned char* validBytes, double maxZErr, unsigned* numBytes);
unsigned lo_encodeForVersion(const void* data, int version, unsigned dt, int nDepth, int nCols, int nRows, int nBands,
                             int nMasks, const unsigned char* validBytes, double maxZErr, unsigned char* out, unsigned outSize,
                             unsigned* nWritten);
unsigned lo_getBlobInfo(const unsigned char* blob, unsigned blobSize, unsigned* infoArray, double* dataRangeArray,
                        int infoArraySize, int dataRangeArraySize);
unsigned lo_getDataRanges(const unsigned char* blob, unsigned blobSize, int nDepth, int nBands, double* mins, double* maxs);
unsigned lo_decode(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                   int nCols, int nRows, int nBands, unsigned dt, void* data);
unsigned lo_decodeToDouble(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                           int nCols, int nRows, int nBands, double* data);

/* the _4D entry points: per-band noData values (Lerc_c_api.h:295-380; Lerc.cpp:1241-1374, :1378-1618, :1046-1076) */
unsigned lo_computeCompressedSize_4D(const void* data, unsigned dt, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                                     const unsigned char* validBytes, double maxZErr, unsigned* numBytes,
                                     const unsigned char* usesNoData, const double* noDataValues);
unsigned lo_encode_4D(const void* data, unsigned dt, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                      const unsigned char* validBytes, double maxZErr, unsigned char* out, unsigned outSize, unsigned* nWritten,
                      const unsigned char* usesNoData, const double* noDataValues);
unsigned lo_decode_4D(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                      int nCols, int nRows, int nBands, unsigned dt, void* data, unsigned char* usesNoData, double* noDataValues);
unsigned lo_decodeToDouble_4D(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                              int nCols, int nRows, int nBands, double* data, unsigned char* usesNoData, double* noDataValues);

/* building blocks exposed for unit tests */
uint32_t lo_fletcher32(const uint8_t* bytes, int len);
size_t   lo_rle_size(const uint8_t* src, size_t n);
size_t   lo_rle_encode(const uint8_t* src, size_t n, uint8_t* dst);           /* returns bytes written */
int      lo_rle_decode(const uint8_t* src, size_t srcLen, uint8_t* dst, size_t dstLen);
int      lo_fpl_encoder(int on);   /* test switch: 0 = float maxZError 0 is written without the FPL codec; returns the previous setting */
int      lo_huffman_lengths(const int* histo, int n, uint16_t* lenOut, uint32_t* codeOut); /* 1 on success */

#ifdef __cplusplus
}
#endif
#endif
