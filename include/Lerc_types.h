/*
 * Lerc_types.h -- integer codes used by the Lerc C API (lerc_b200 drop-in).
 *
 * Same names and values as the reference's src/LercLib/include/Lerc_types.h:11-56, so C++ callers
 * that include <Lerc_types.h> and use LercNS::ErrCode / DataType / InfoArrOrder / DataRangeArrOrder
 * compile unchanged against this library.  Callers from other languages only ever see the integers.
 */
#pragma once

namespace LercNS
{
  /* lerc_status values returned by every API function (reference Lerc_types.h:11-20) */
  enum class ErrCode : int
  { Ok = 0, Failed, WrongParam, BufferTooSmall, NaN, HasNoData, DimensionsTooLarge };

  /* pixel types, the `dataType` argument (reference Lerc_types.h:22-32) */
  enum class DataType : int
  { dt_char = 0, dt_uchar, dt_short, dt_ushort, dt_int, dt_uint, dt_float, dt_double };

  /* slots of lerc_getBlobInfo()'s infoArray (reference Lerc_types.h:34-48); nDim is the old name of nDepth */
  enum class InfoArrOrder : int
  { version = 0, dataType, nDim, nCols, nRows, nBands, nValidPixels, blobSize, nMasks, nDepth, nUsesNoDataValue, _last };

  /* slots of lerc_getBlobInfo()'s dataRangeArray (reference Lerc_types.h:50-56) */
  enum class DataRangeArrOrder : int
  { zMin = 0, zMax, maxZErrUsed, _last };
}
