// lerc_lookback.cuh -- decoupled look-back over per-tile aggregates (single-pass chained scan), two levels.
//
// Every tile (stream chunk) publishes its aggregate in st[tile] as soon as it is known; one warp per tile turns it into
// the exclusive prefix over all earlier tiles: round 1 adds the aggregates of the predecessors inside the tile's group of
// 32, round 2 the aggregates of whole groups (accumulated atomically by the publishers in gacc[g]; the warp that looks back
// for a group's last tile adds the group's inclusive prefix in gs[g]), so a look-back is two or three dependent rounds of
// loads however many tiles are in flight, and depends on nothing but the publications.  Words are {2-bit state, 62-bit
// value}: 0 = not there yet, ST_A = aggregate, ST_P = inclusive prefix.  The polling is warp-uniform (one round of loads,
// a vote, a sleep), which keeps it cheap in issue slots and lets tools/cusim schedule the CTA's other threads.
#pragma once
#include "lerc_device.cuh"

namespace lerc {

constexpr unsigned long long LB_A = 1ull << 62, LB_P = 2ull << 62, LB_VAL = (1ull << 62) - 1;
constexpr unsigned long long LB_ONE = 1ull << 56, LB_SUM = LB_ONE - 1;     // group accumulators: bits 56..61 count the tiles, bits 0..55 sum their values

// aggregate of `tile`, published before the look-back (by any thread): the tile's own word and its share of the group's
// accumulator, so that a group's aggregate is complete as soon as its 32 tiles are published, whoever looks back for them
__device__ __forceinline__ void lookbackPublish(volatile unsigned long long* st, unsigned long long* gacc, long long tile, unsigned long long value) {
  st[tile] = (tile == 0 ? LB_P : LB_A) | value;
  atomicAdd(&gacc[tile >> 5], LB_ONE + value);
}

// Exclusive prefix of `tile` (whole warp, converged).  Publishes the tile's inclusive prefix and, for the last tile of a
// group, the group's inclusive prefix in gs[].
__device__ __forceinline__ unsigned long long lookbackExclusive(volatile unsigned long long* st, const unsigned long long* gacc, volatile unsigned long long* gs,
                                                                long long tile, unsigned long long value, int lane) {
  unsigned long long excl = 0;
  const int l = (int)(tile & 31);
  const long long g = tile >> 5;
  bool needGroups = g > 0;
  {  // round 1: the predecessors inside the tile's group of 32
    const long long idx = tile - 1 - lane;
    const bool in = lane < l;
    unsigned long long s = 0;
    for (;;) {
      if (in && (s >> 62) == 0) s = st[idx];
      if (__all_sync(FULL, !in || (s >> 62) != 0)) break;
      __nanosleep(100);
    }
    const unsigned isP = __ballot_sync(FULL, in && (s >> 62) == 2);
    const int firstP = isP ? __ffs(isP) - 1 : 32;
    unsigned long long contrib = (in && lane <= firstP) ? (s & LB_VAL) : 0;
#pragma unroll
    for (int m = 16; m; m >>= 1) contrib += __shfl_xor_sync(FULL, contrib, m);
    excl = contrib;
    if (isP) needGroups = false;                                    // an inclusive prefix inside the group: excl is already global
  }
  if (needGroups) {  // round 2: whole groups, 32 at a time: a group's inclusive prefix if somebody has published it, else its full accumulator
    long long base = g - 1;
    for (;;) {
      const long long idx = base - lane;
      unsigned long long s = idx >= 0 ? 0ull : LB_P;                // virtual groups before 0: prefix 0
      for (;;) {
        if ((s >> 62) == 0) {
          s = gs[idx];
          if ((s >> 62) == 0) {
            const unsigned long long acc = *(volatile const unsigned long long*)&gacc[idx];
            if ((acc >> 56) == 32) s = LB_A | (acc & LB_SUM);
          }
        }
        if (__all_sync(FULL, (s >> 62) != 0)) break;
        __nanosleep(100);
      }
      const unsigned isP = __ballot_sync(FULL, (s >> 62) == 2);
      const int firstP = isP ? __ffs(isP) - 1 : 32;
      unsigned long long contrib = (lane <= firstP && idx >= 0) ? (s & LB_VAL) : 0;
#pragma unroll
      for (int m = 16; m; m >>= 1) contrib += __shfl_xor_sync(FULL, contrib, m);
      excl += contrib;
      if (isP) break;
      base -= 32;
    }
  }
  if (lane == 0) {
    if (tile > 0) st[tile] = LB_P | (excl + value);
    if (l == 31) gs[g] = LB_P | (excl + value);                     // inclusive prefix of the whole group
  }
  return excl;
}

}  // namespace lerc
