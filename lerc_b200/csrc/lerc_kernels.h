// lerc_kernels.h -- host-callable launchers of the kernels in lerc_mask.cu / lerc_encode.cu /
// lerc_decode.cu.  Everything runs on ctx->stream; nothing here synchronises.
#pragma once
#include "lerc_internal.h"
#include <algorithm>

namespace lerc {

enum { MASKF_MODIFIED = 1, MASKF_MIXED_NAN = 2 };

// ---- lerc_mask.cu ----
template <class T>
void launchMaskBuild(Context* ctx, const void* dData, const uint8_t* dBytes, long long nPix, int nDepth, uint8_t* dBits, int* dCounters);
void launchBitsToBytes(Context* ctx, const uint8_t* dBits, long long nPix, uint8_t* dBytes);
void launchBitsDiffer(Context* ctx, const uint8_t* a, const uint8_t* b, long long nPix, int* dFlag);
void launchChunkValidCounts(Context* ctx, const uint8_t* dBits, long long nPix, int nChunks, uint32_t* dCounts);
void launchRleEncode(Context* ctx, const uint8_t* dSrc, long long n, uint8_t* dDst, uint32_t* dSize);
void launchRleDecode(Context* ctx, const uint8_t* dSrc, long long srcLen, uint8_t* dDst, long long dstLen, int* dStatus);
void launchFletcher(Context* ctx, const uint8_t* dRegion, long long len, unsigned long long* dAcc, uint8_t* dStoreAt,
                    uint32_t expect, int* dStatus);

// exclusive prefix sum of n uint32 values (n + 1 outputs, the last one is the total); lerc_encode.cu
void exclusiveScanU32(Context* ctx, const uint32_t* dIn, uint32_t* dOut, size_t n);
void exclusiveScanU64(Context* ctx, const unsigned long long* dIn, unsigned long long* dOut, size_t n);

}  // namespace lerc
