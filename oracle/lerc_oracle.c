/*
 * lerc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See lerc_oracle.h for scope and parity status.
 *
 * Untyped building blocks of the Lerc2 restatement (header, Fletcher-32, RLE, fixed-width bit
 * stuffing, canonical Huffman) followed by the 8 typed instantiations of lerc_oracle_typed.inc and
 * the lo_* API.  Every function names the reference file:line whose behaviour it restates
 * (paths relative to /root/reference/src/LercLib).
 */
#include "lerc_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>
#include <float.h>

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

static const int kTypeSize[8] = {1, 1, 2, 2, 4, 4, 4, 8};

/* ------------------------------------------------------------------------------------------- */
/* header                                                                        Lerc2.cpp:710-917 */

typedef struct {
  int version;
  u32 checksum;
  int nRows, nCols, nDepth, numValid, mbSize, blobSize, dt, nBlobsMore;
  u8 passNoData, isInt, res3, res4;
  double maxZErr, zMin, zMax, noData, noDataOrig;
} lo_hdr;

static int hdr_bytes(int v) {
  return 6 + 4 + (v >= 3 ? 4 : 0) + 4 * (v >= 4 ? 7 : 6) + (v >= 6 ? 8 : 0) + 8 * (v >= 6 ? 5 : 3);
}

static void put_i32(u8** p, int v) { memcpy(*p, &v, 4); *p += 4; }
static void put_f64(u8** p, double v) { memcpy(*p, &v, 8); *p += 8; }
static int get_i32(const u8** p) { int v; memcpy(&v, *p, 4); *p += 4; return v; }
static double get_f64(const u8** p) { double v; memcpy(&v, *p, 8); *p += 8; return v; }

/* Lerc2.cpp:724-786.  The checksum slot is left 0; finish_blob() fills it. */
static void hdr_write(u8* p, const lo_hdr* h) {
  memcpy(p, "Lerc2 ", 6); p += 6;
  put_i32(&p, h->version);
  if (h->version >= 3) put_i32(&p, 0);
  put_i32(&p, h->nRows); put_i32(&p, h->nCols);
  if (h->version >= 4) put_i32(&p, h->nDepth);
  put_i32(&p, h->numValid); put_i32(&p, h->mbSize); put_i32(&p, h->blobSize); put_i32(&p, h->dt);
  if (h->version >= 6) {
    put_i32(&p, h->nBlobsMore);
    *p++ = h->passNoData; *p++ = h->isInt; *p++ = h->res3; *p++ = h->res4;
  }
  put_f64(&p, h->maxZErr); put_f64(&p, h->zMin); put_f64(&p, h->zMax);
  if (h->version >= 6) { put_f64(&p, h->noData); put_f64(&p, h->noDataOrig); }
}

/* Lerc2.cpp:790-917 incl. the dimension guards at :877-911.  Returns header length or 0. */
static int hdr_read(const u8* p, size_t avail, lo_hdr* h) {
  memset(h, 0, sizeof *h);
  if (avail < 10 || memcmp(p, "Lerc2 ", 6)) return 0;
  const u8* q = p + 6;
  h->version = get_i32(&q);
  if (h->version < 0 || h->version > 6) return 0;
  int need = hdr_bytes(h->version);
  if (avail < (size_t)need) return 0;
  if (h->version >= 3) h->checksum = (u32)get_i32(&q);
  h->nRows = get_i32(&q); h->nCols = get_i32(&q);
  h->nDepth = h->version >= 4 ? get_i32(&q) : 1;
  h->numValid = get_i32(&q); h->mbSize = get_i32(&q); h->blobSize = get_i32(&q);
  int dt = get_i32(&q);
  if (h->version >= 6) {
    h->nBlobsMore = get_i32(&q);
    h->passNoData = *q++; h->isInt = *q++; h->res3 = *q++; h->res4 = *q++;
  }
  h->maxZErr = get_f64(&q); h->zMin = get_f64(&q); h->zMax = get_f64(&q);
  if (h->version >= 6) { h->noData = get_f64(&q); h->noDataOrig = get_f64(&q); }
  if (h->nRows <= 0 || h->nCols <= 0 || h->nDepth <= 0 || h->numValid < 0 || h->mbSize <= 0 || h->blobSize <= 0 ||
      dt < 0 || dt > LO_DOUBLE)
    return 0;
  h->dt = dt;
  u64 nPix = (u64)h->nRows * (u64)h->nCols, lim = (u64)INT_MAX, bpp = (u64)kTypeSize[dt];
  if (nPix > lim || (u64)h->numValid > nPix) return 0;
  if (h->mbSize > 32 || bpp * (u64)h->nDepth > lim || bpp * (u64)h->nDepth * nPix > lim) return 0;
  return need;
}

/* ------------------------------------------------------------------------------------------- */
/* Fletcher-32 over big-endian 16-bit words                                    Lerc2.cpp:1037-1064 */

uint32_t lo_fletcher32(const uint8_t* b, int len) {
  u32 s1 = 0xffff, s2 = 0xffff;
  int words = len / 2;
  while (words > 0) {
    int chunk = words > 359 ? 359 : words;   /* 359 words keep both sums below 2^32 before folding */
    words -= chunk;
    for (int i = 0; i < chunk; i++, b += 2) {
      s1 += ((u32)b[0] << 8) + b[1];
      s2 += s1;
    }
    s1 = (s1 & 0xffff) + (s1 >> 16);
    s2 = (s2 & 0xffff) + (s2 >> 16);
  }
  if (len & 1) { s1 += (u32)b[0] << 8; s2 += s1; }
  s1 = (s1 & 0xffff) + (s1 >> 16);
  s2 = (s2 & 0xffff) + (s2 >> 16);
  return (s2 << 16) | s1;
}

/* ------------------------------------------------------------------------------------------- */
/* byte RLE of the bit mask                                                         RLE.cpp:32-331 */
/* Restated as a run tokenizer: a maximal run of equal bytes becomes a "repeat" token iff it is at
 * least 5 long and starts more than 5 bytes before the end of the array (RLE.cpp:74-79); everything
 * else is literal.  Both token kinds are cut into pieces of at most 32767 (RLE.cpp:98-107).
 * Token = int16 count (LE) then: count>0 -> that many literal bytes; count<0 -> one byte repeated
 * -count times; -32768 terminates (RLE.cpp:250). */

typedef struct { u8* dst; size_t n; } rle_sink;

static void rle_put(rle_sink* s, int count, const u8* bytes, size_t nBytes) {
  if (s->dst) {
    int16_t c = (int16_t)count;
    s->dst[s->n] = (u8)(c & 0xff);
    s->dst[s->n + 1] = (u8)((c >> 8) & 0xff);
    if (nBytes) memcpy(s->dst + s->n + 2, bytes, nBytes);
  }
  s->n += 2 + nBytes;
}

static void rle_literals(rle_sink* s, const u8* src, size_t a, size_t b) {
  while (a < b) {
    size_t c = b - a > 32767 ? 32767 : b - a;
    rle_put(s, (int)c, src + a, c);
    a += c;
  }
}

static size_t rle_run(const u8* src, size_t n, u8* dst) {
  rle_sink s = {dst, 0};
  if (!src || n == 0) return 0;
  size_t pos = 0, lit = 0;
  while (pos < n) {
    size_t run = 1;
    while (pos + run < n && src[pos + run] == src[pos]) run++;
    if (run >= 5 && pos + 5 < n) {
      rle_literals(&s, src, lit, pos);
      size_t left = run;
      while (left) {
        size_t c = left > 32767 ? 32767 : left;
        rle_put(&s, -(int)c, src + pos, 1);
        left -= c;
      }
      lit = pos + run;
    }
    pos += run;
  }
  rle_literals(&s, src, lit, n);
  rle_put(&s, -32768, NULL, 0);
  return s.n;
}

size_t lo_rle_size(const uint8_t* src, size_t n) { return rle_run(src, n, NULL); }
size_t lo_rle_encode(const uint8_t* src, size_t n, uint8_t* dst) { return rle_run(src, n, dst); }

/* RLE.cpp:298-331.  1 on success. */
int lo_rle_decode(const uint8_t* src, size_t srcLen, uint8_t* dst, size_t dstLen) {
  if (!src || !dst || srcLen < 2) return 0;
  size_t ip = 0, op = 0;
  for (;;) {
    if (ip + 2 > srcLen) return 0;
    int16_t c = (int16_t)(src[ip] | (src[ip + 1] << 8));
    ip += 2;
    if (c == -32768) return 1;
    size_t cnt = (size_t)(c <= 0 ? -c : c), take = c > 0 ? cnt : 1;
    if (ip + take + 2 > srcLen || op + cnt > dstLen) return 0;
    if (c > 0) memcpy(dst + op, src + ip, cnt);
    else memset(dst + op, src[ip], cnt);
    ip += take; op += cnt;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* fixed-width bit stuffing, v3+ layout                                    BitStuffer2.cpp:35-287 */

static int bit_length(u32 v) { int n = 0; while (n < 32 && (v >> n)) n++; return n; }
static int count_field_bytes(u32 n) { return n < 256 ? 1 : (n < 65536 ? 2 : 4); }
static size_t packed_bytes(u32 n, int nb) { return (size_t)(((u64)n * (u64)nb + 7) >> 3); }

/* value i occupies stream bits [i*nb, (i+1)*nb), bit k of the stream = bit (k&7) of byte k>>3.
 * This equals the reference's LE-uint32, LSB-first words with the unused tail bytes dropped
 * (BitStuffer2.cpp:432-472). */
static void pack_lsb(u8* dst, const u32* v, u32 n, int nb) {
  size_t len = packed_bytes(n, nb);
  memset(dst, 0, len);
  u64 bit = 0;
  for (u32 i = 0; i < n; i++, bit += (u64)nb) {
    u64 x = (u64)v[i] << (bit & 7);
    for (size_t k = bit >> 3; x; k++, x >>= 8) dst[k] |= (u8)x;
  }
}

static void unpack_lsb(const u8* src, u32* v, u32 n, int nb) {
  size_t len = packed_bytes(n, nb);
  u64 bit = 0;
  u32 mask = nb == 32 ? 0xffffffffu : ((1u << nb) - 1);
  for (u32 i = 0; i < n; i++, bit += (u64)nb) {
    u64 x = 0;
    size_t k0 = bit >> 3;
    for (int k = 0; k < 5 && k0 + k < len; k++) x |= (u64)src[k0 + k] << (8 * k);
    v[i] = (u32)(x >> (bit & 7)) & mask;
  }
}

/* Lerc2 v2 bit stuffing (BitStuffer2.cpp:292-425): values MSB-first inside little-endian uint32 words; the unused low
 * bytes of the last word are dropped by shifting that word down (so its used bytes come first). */
static u32 tail_bytes_not_needed(u32 n, int nb) {             /* BitStuffer2.h:127-132 */
  int bitsTail = (int)(((u64)n * (u64)nb) & 31), bytesTail = (bitsTail + 7) >> 3;
  return bytesTail > 0 ? (u32)(4 - bytesTail) : 0;
}
static void pack_msb_v2(u8* dst, const u32* v, u32 n, int nb) {
  size_t nWords = (size_t)(((u64)n * (u64)nb + 31) / 32);
  u32* w = (u32*)calloc(nWords + 1, sizeof(u32));
  size_t k = 0; int bitPos = 0;
  for (u32 i = 0; i < n; i++) {
    if (32 - bitPos >= nb) {
      w[k] |= v[i] << (32 - bitPos - nb);
      bitPos += nb;
      if (bitPos == 32) { bitPos = 0; k++; }
    } else {
      int r = nb - (32 - bitPos);
      w[k] |= v[i] >> r; k++;
      w[k] |= v[i] << (32 - r);
      bitPos = r;
    }
  }
  u32 drop = tail_bytes_not_needed(n, nb);
  if (nWords) w[nWords - 1] >>= 8 * drop;
  memcpy(dst, w, nWords * 4 - drop);            /* == packed_bytes(n, nb) */
  free(w);
}
static void unpack_msb_v2(const u8* src, u32* v, u32 n, int nb) {
  size_t nWords = (size_t)(((u64)n * (u64)nb + 31) / 32), len = packed_bytes(n, nb);
  u32* w = (u32*)calloc(nWords + 1, sizeof(u32));
  memcpy(w, src, len);
  u32 drop = tail_bytes_not_needed(n, nb);
  if (nWords) w[nWords - 1] <<= 8 * drop;
  size_t k = 0; int bitPos = 0;
  for (u32 i = 0; i < n; i++) {
    if (32 - bitPos >= nb) {
      v[i] = (w[k] << bitPos) >> (32 - nb);
      bitPos += nb;
      if (bitPos == 32) { bitPos = 0; k++; }
    } else {
      u32 hi = (w[k] << bitPos) >> (32 - nb); k++;
      bitPos -= 32 - nb;
      v[i] = hi | (w[k] >> (32 - bitPos));
    }
  }
  free(w);
}
static void pack_bits(u8* dst, const u32* v, u32 n, int nb, int version) { if (version >= 3) pack_lsb(dst, v, n, nb); else pack_msb_v2(dst, v, n, nb); }
static void unpack_bits(const u8* src, u32* v, u32 n, int nb, int version) { if (version >= 3) unpack_lsb(src, v, n, nb); else unpack_msb_v2(src, v, n, nb); }

static u8* put_count(u8* p, u32 n, int nBytes) {
  if (nBytes == 1) *p = (u8)n;
  else if (nBytes == 2) { uint16_t s = (uint16_t)n; memcpy(p, &s, 2); }
  else memcpy(p, &n, 4);
  return p + nBytes;
}

static u32 simple_size(u32 n, u32 maxElem) {   /* BitStuffer2.h:68-74 */
  return 1 + (u32)count_field_bytes(n) + (u32)packed_bytes(n, bit_length(maxElem));
}

/* BitStuffer2.cpp:35-75 */
static u8* encode_simple(u8* p, const u32* v, u32 n, int version) {
  u32 mx = 0;
  for (u32 i = 0; i < n; i++) if (v[i] > mx) mx = v[i];
  int nb = bit_length(mx), cb = count_field_bytes(n);
  *p++ = (u8)(nb | ((cb == 4 ? 0 : 3 - cb) << 6));
  p = put_count(p, n, cb);
  if (nb > 0) { pack_bits(p, v, n, nb, version); p += packed_bytes(n, nb); }
  return p;
}

static int cmp_u32(const void* a, const void* b) {
  u32 x = *(const u32*)a, y = *(const u32*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* distinct sorted values of v[] into lut[] (including the leading one); returns their count */
static u32 distinct_sorted(const u32* v, u32 n, u32* lut) {
  memcpy(lut, v, n * sizeof(u32));
  qsort(lut, n, sizeof(u32), cmp_u32);
  u32 m = 0;
  for (u32 i = 0; i < n; i++) if (i == 0 || lut[i] != lut[i - 1]) lut[m++] = lut[i];
  return m;
}

/* BitStuffer2.cpp:262-287: size of the cheaper of {simple, LUT}; *useLut says which. */
static u32 lut_or_simple_size(const u32* v, u32 n, u32* scratch, int* useLut) {
  u32 m = distinct_sorted(v, n, scratch);
  u32 nLut = m - 1;
  int nb = bit_length(scratch[m - 1]), nbIdx = bit_length(nLut), cb = count_field_bytes(n);
  u32 simple = 1 + (u32)cb + (u32)packed_bytes(n, nb);
  u32 lut = 1 + (u32)cb + 1 + (u32)packed_bytes(nLut, nb) + (u32)packed_bytes(n, nbIdx);
  *useLut = lut < simple;
  return lut < simple ? lut : simple;
}

/* BitStuffer2.cpp:79-153.  v[] must contain a 0 (the block minimum).  NULL on failure. */
static u8* encode_lut(u8* p, const u32* v, u32 n, u32* scratch, u32* idxScratch, int version) {
  u32 m = distinct_sorted(v, n, scratch);
  if (m < 2 || m > 255 || scratch[0] != 0) return NULL;
  u32 nLut = m - 1;
  int nb = bit_length(scratch[m - 1]), nbIdx = bit_length(nLut), cb = count_field_bytes(n);
  if (nb <= 0 || nb >= 32) return NULL;
  for (u32 i = 0; i < n; i++) {     /* rank of v[i] among the distinct values */
    u32 lo = 0, hi = m - 1;
    while (lo < hi) { u32 mid = (lo + hi) / 2; if (scratch[mid] < v[i]) lo = mid + 1; else hi = mid; }
    idxScratch[i] = lo;
  }
  *p++ = (u8)(nb | (1 << 5) | ((cb == 4 ? 0 : 3 - cb) << 6));
  p = put_count(p, n, cb);
  *p++ = (u8)(nLut + 1);
  pack_bits(p, scratch + 1, nLut, nb, version); p += packed_bytes(nLut, nb);
  pack_bits(p, idxScratch, n, nbIdx, version);  p += packed_bytes(n, nbIdx);
  return p;
}

/* BitStuffer2.cpp:159-258 (both bit orders).  Returns bytes consumed, 0 on malformed input.
 * v[] must hold maxCount entries. */
static size_t decode_bitstuffed(const u8* p, size_t avail, u32* v, u32 maxCount, u32* nOut, int version) {
  const u8* p0 = p;
  if (avail < 1) return 0;
  u8 b = *p++; avail--;
  int code = b >> 6, cb = code == 0 ? 4 : 3 - code, lut = (b >> 5) & 1, nb = b & 31;
  if (cb <= 0 || avail < (size_t)cb) return 0;
  u32 n = 0;
  if (cb == 1) n = *p; else if (cb == 2) { uint16_t s; memcpy(&s, p, 2); n = s; } else memcpy(&n, p, 4);
  p += cb; avail -= cb;
  if (n > maxCount) return 0;
  if (!lut) {
    if (nb > 0) {
      if (n == 0) return 0;
      size_t len = packed_bytes(n, nb);
      if (avail < len) return 0;
      unpack_bits(p, v, n, nb, version); p += len;
    } else memset(v, 0, n * sizeof(u32));
  } else {
    if (nb == 0 || avail < 1) return 0;
    int nLut = (int)*p++ - 1; avail--;
    if (nLut < 1 || n == 0) return 0;
    u32 table[256];
    size_t len = packed_bytes((u32)nLut, nb);
    if (avail < len) return 0;
    table[0] = 0;
    unpack_bits(p, table + 1, (u32)nLut, nb, version); p += len; avail -= len;
    int nbIdx = bit_length((u32)nLut);
    len = packed_bytes(n, nbIdx);
    if (avail < len) return 0;
    unpack_bits(p, v, n, nbIdx, version); p += len;
    for (u32 i = 0; i < n; i++) { if (v[i] > (u32)nLut) return 0; v[i] = table[v[i]]; }
  }
  *nOut = n;
  return (size_t)(p - p0);
}

/* ------------------------------------------------------------------------------------------- */
/* canonical Huffman over a 256-bin histogram                                  Huffman.cpp:35-572 */

typedef struct { int weight; int leaf; int kid0, kid1; } hnode;   /* weight = -count (Huffman.h:90) */

/* The reference keeps its nodes in std::priority_queue<Node, vector<Node>, less<Node>>
 * (Huffman.cpp:40).  Which of several equal-weight nodes surfaces first is decided by the binary
 * heap's sift order, and that decides the code lengths, so the two libstdc++ sift routines
 * (bits/stl_heap.h: __push_heap, __adjust_heap) are restated here on an index heap. */
static void heap_sift_up(int* heap, const hnode* nd, int hole, int top, int val) {
  int parent = (hole - 1) / 2;
  while (hole > top && nd[heap[parent]].weight < nd[val].weight) {
    heap[hole] = heap[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  heap[hole] = val;
}

static void heap_push(int* heap, int* n, const hnode* nd, int val) {
  heap[*n] = val; (*n)++;
  heap_sift_up(heap, nd, *n - 1, 0, val);
}

static int heap_pop(int* heap, int* n, const hnode* nd) {
  int top = heap[0];
  int len = --(*n);
  if (len > 0) {
    int val = heap[len], hole = 0, kid = 0;
    while (kid < (len - 1) / 2) {
      kid = 2 * (kid + 1);
      if (nd[heap[kid]].weight < nd[heap[kid - 1]].weight) kid--;
      heap[hole] = heap[kid];
      hole = kid;
    }
    if ((len & 1) == 0 && kid == (len - 2) / 2) {
      kid = 2 * (kid + 1);
      heap[hole] = heap[kid - 1];
      hole = kid - 1;
    }
    heap_sift_up(heap, nd, hole, 0, val);
  }
  return top;
}

static int assign_depths(const hnode* nd, int root, int depth, uint16_t* len) {
  if (nd[root].leaf >= 0) { len[nd[root].leaf] = (uint16_t)depth; return 1; }
  if (depth == 32) return 0;                               /* Huffman.h:91 */
  return assign_depths(nd, nd[root].kid0, depth + 1, len) && assign_depths(nd, nd[root].kid1, depth + 1, len);
}

typedef struct { int key; int sym; } canon_key;
static int cmp_canon(const void* a, const void* b) {
  int x = ((const canon_key*)a)->key, y = ((const canon_key*)b)->key;
  return x > y ? -1 : (x < y ? 1 : 0);
}

/* Huffman.cpp:35-81 + :541-572.  1 on success (>= 2 used symbols, no code longer than 32). */
int lo_huffman_lengths(const int* histo, int n, uint16_t* len, uint32_t* code) {
  hnode nd[512]; int heap[256], hn = 0, nn = 0;
  if (n > 256) return 0;
  for (int i = 0; i < n; i++) { len[i] = 0; code[i] = 0; }
  for (int i = 0; i < n; i++)
    if (histo[i] > 0) { nd[nn].weight = -histo[i]; nd[nn].leaf = i; nd[nn].kid0 = nd[nn].kid1 = -1; heap_push(heap, &hn, nd, nn); nn++; }
  if (hn < 2) return 0;
  while (hn > 1) {
    int a = heap_pop(heap, &hn, nd), b = heap_pop(heap, &hn, nd);
    nd[nn].weight = nd[a].weight + nd[b].weight; nd[nn].leaf = -1; nd[nn].kid0 = a; nd[nn].kid1 = b;
    heap_push(heap, &hn, nd, nn); nn++;
  }
  if (!assign_depths(nd, heap[0], 0, len)) return 0;
  /* canonical codes: longest first, smaller symbol first within a length, code shrinks by >> */
  canon_key keys[256]; int nk = 0;
  for (int i = 0; i < n; i++) if (len[i] > 0) { keys[nk].key = len[i] * n - i; keys[nk].sym = i; nk++; }
  qsort(keys, nk, sizeof(canon_key), cmp_canon);
  int curLen = len[keys[0].sym]; u32 c = 0;
  for (int k = 0; k < nk; k++) {
    int s = keys[k].sym, d = curLen - len[s];
    c >>= d; curLen -= d;
    code[s] = c++;
  }
  return 1;
}

static int wrap(int i, int size) { return i < size ? i : i - size; }

/* Huffman.cpp:383-438: the stored index range, possibly wrapping past `size`. */
static int huff_range(const uint16_t* len, int size, int* i0, int* i1, int* maxLen) {
  int a = 0, b = size - 1;
  while (a < size && len[a] == 0) a++;
  while (b >= 0 && len[b] == 0) b--;
  if (b + 1 <= a) return 0;
  *i0 = a; *i1 = b + 1;
  int bestStart = 0, bestLen = 0, j = 0;
  while (j < size) {
    while (j < size && len[j] > 0) j++;
    int k0 = j;
    while (j < size && len[j] == 0) j++;
    if (j - k0 > bestLen) { bestStart = k0; bestLen = j - k0; }
  }
  if (size - bestLen < *i1 - *i0) { *i0 = bestStart + bestLen; *i1 = bestStart + size; }
  if (*i1 <= *i0) return 0;
  int mx = 0;
  for (int i = *i0; i < *i1; i++) { int l = len[wrap(i, size)]; if (l > mx) mx = l; }
  if (mx <= 0 || mx > 32) return 0;
  *maxLen = mx;
  return 1;
}

/* Huffman.cpp:357-379 */
static int huff_table_bytes(const uint16_t* len, int size, int* nBytes) {
  int i0, i1, maxLen;
  if (!huff_range(len, size, &i0, &i1, &maxLen)) return 0;
  int sum = 0;
  for (int i = i0; i < i1; i++) sum += len[wrap(i, size)];
  *nBytes = 16 + (int)simple_size((u32)(i1 - i0), (u32)maxLen) + 4 * ((((sum + 7) >> 3) + 3) >> 2);
  return 1;
}

/* Huffman.cpp:85-111; total bits >= 2^31 is reported as "not available" (see DESIGN.md). */
static int huff_total_bytes(const int* histo, const uint16_t* len, int size, int* nBytes) {
  int tb;
  if (!huff_table_bytes(len, size, &tb)) return 0;
  int64_t bits = 0, elems = 0;
  for (int i = 0; i < size; i++) if (histo[i] > 0) { bits += (int64_t)histo[i] * len[i]; elems += histo[i]; }
  if (elems == 0 || bits >= ((int64_t)1 << 31)) return 0;
  int64_t words = ((((bits + 7) >> 3) + 3) >> 2) + 1;
  int64_t total = tb + 4 * words;
  if (total > INT_MAX) return 0;
  *nBytes = (int)total;
  return 1;
}

/* MSB-first bit writer into little-endian uint32 words                        Huffman.h:218-255 */
typedef struct { u8* base; u64 bitPos; } msb_writer;
static void msb_put(msb_writer* w, u32 val, int nBits) {
  for (int k = nBits - 1; k >= 0; k--, w->bitPos++) {
    if ((val >> k) & 1) {
      u64 word = w->bitPos >> 5; int bit = 31 - (int)(w->bitPos & 31);   /* bit index inside the LE word */
      w->base[word * 4 + (bit >> 3)] |= (u8)(1u << (bit & 7));
    }
  }
}
static int msb_get(const u8* base, u64 bitPos) {
  u64 word = bitPos >> 5; int bit = 31 - (int)(bitPos & 31);
  return (base[word * 4 + (bit >> 3)] >> (bit & 7)) & 1;
}

/* Huffman.cpp:126-166 + :442-467.  Destination must be zero-filled.  Returns bytes written. */
static size_t huff_write_table(u8* p, const uint16_t* len, const u32* code, int size, int version) {
  u8* p0 = p;
  int i0, i1, maxLen;
  if (!huff_range(len, size, &i0, &i1, &maxLen)) return 0;
  put_i32(&p, 4); put_i32(&p, size); put_i32(&p, i0); put_i32(&p, i1);
  u32 lens[512];
  for (int i = i0; i < i1; i++) lens[i - i0] = len[wrap(i, size)];
  p = encode_simple(p, lens, (u32)(i1 - i0), version);
  msb_writer w = {p, 0};
  for (int i = i0; i < i1; i++) { int k = wrap(i, size); if (len[k] > 0) msb_put(&w, code[k], len[k]); }
  p += 4 * ((w.bitPos + 31) >> 5);
  return (size_t)(p - p0);
}

/* Huffman.cpp:170-234 + :471-537.  Returns bytes consumed or 0. */
static size_t huff_read_table(const u8* p, size_t avail, uint16_t* len, u32* code, int* sizeOut, int version) {
  const u8* p0 = p;
  if (avail < 16) return 0;
  int ver = get_i32(&p), size = get_i32(&p), i0 = get_i32(&p), i1 = get_i32(&p);
  avail -= 16;
  if (ver < 2 || i0 >= i1 || i0 < 0 || size < 0 || size > 256) return 0;
  if (wrap(i0, size) >= size || wrap(i1 - 1, size) >= size || i1 - i0 > 512) return 0;
  u32 lens[512], n = 0;
  size_t used = decode_bitstuffed(p, avail, lens, (u32)(i1 - i0), &n, version);
  if (!used || n != (u32)(i1 - i0)) return 0;
  p += used; avail -= used;
  for (int i = 0; i < size; i++) { len[i] = 0; code[i] = 0; }
  u64 bits = 0;
  for (int i = i0; i < i1; i++) {
    int k = wrap(i, size);
    if (lens[i - i0] > 32) return 0;
    len[k] = (uint16_t)lens[i - i0];
    bits += len[k];
  }
  size_t bytes = 4 * (size_t)((bits + 31) >> 5);
  if (avail < bytes) return 0;
  u64 pos = 0;
  for (int i = i0; i < i1; i++) {
    int k = wrap(i, size);
    u32 c = 0;
    for (int b = 0; b < len[k]; b++) c = (c << 1) | (u32)msb_get(p, pos++);
    code[k] = c;
  }
  p += bytes;
  *sizeOut = size;
  return (size_t)(p - p0);
}

/* prefix-code decoder as a binary trie over the explicit (length, code) pairs of the table */
typedef struct { int16_t kid[2]; int16_t sym; } trie_node;
typedef struct { trie_node n[1024]; int used; } trie;

static int trie_build(trie* t, const uint16_t* len, const u32* code, int size) {
  t->used = 1; t->n[0].kid[0] = t->n[0].kid[1] = -1; t->n[0].sym = -1;
  int any = 0;
  for (int s = 0; s < size; s++) {
    if (!len[s]) continue;
    any = 1;
    int cur = 0;
    for (int b = len[s] - 1; b >= 0; b--) {
      int bit = (code[s] >> b) & 1;
      if (t->n[cur].sym >= 0) return 0;                 /* not prefix free */
      if (t->n[cur].kid[bit] < 0) {
        if (t->used >= 1024) return 0;
        int nn = t->used++;
        t->n[nn].kid[0] = t->n[nn].kid[1] = -1; t->n[nn].sym = -1;
        t->n[cur].kid[bit] = (int16_t)nn;
      }
      cur = t->n[cur].kid[bit];
    }
    if (t->n[cur].kid[0] >= 0 || t->n[cur].kid[1] >= 0 || t->n[cur].sym >= 0) return 0;
    t->n[cur].sym = (int16_t)s;
  }
  return any;
}

/* ------------------------------------------------------------------------------------------- */
/* Lossless float / double codec "FPL" (IEM_DeltaDeltaHuffman), DECODE side.
 * fpl_Lerc2Ext.cpp:725-866 (DecodeHuffmanFlt[Slice]), :133-169 (restoreSequence), :612-722 (restoreCrossBytes / restoreByteOrder);
 * fpl_EsriHuffman.cpp:37-75 (decodePackBits), :453-558 (DecodeHuffman); fpl_UnitTypes.cpp:40-63, :98-111, :140-156, :626-735, :775-888.
 * The whole array is coded (invalid pixels included), one byte plane of the (bit-transformed) values at a time. */
static int fpl_plane_decode(const u8* p, size_t size, u8* out, size_t n) {
  if (size < 1) return 0;
  switch (p[0]) {
    case 1: {                                             /* HUFFMAN_RLE: one value */
      if (size < 6) return 0;
      u32 cnt; memcpy(&cnt, p + 2, 4);
      if ((size_t)cnt != n) return 0;
      memset(out, p[1], n);
      return 1;
    }
    case 2:                                               /* HUFFMAN_NO_ENCODING */
      if (size < 1 + n) return 0;
      memcpy(out, p + 1, n);
      return 1;
    case 3: {                                             /* HUFFMAN_PACKBITS */
      size_t cur = 0, i = 1;
      while (i < size) {
        int b = p[i++];
        if (b <= 127) { size_t c = (size_t)b + 1; if (cur + (size_t)b >= n || i + c > size) return 0; memcpy(out + cur, p + i, c); cur += c; i += c; }
        else { size_t c = (size_t)b - 127 + 1; if (cur + (size_t)b - 127 >= n || i >= size) return 0; memset(out + cur, p[i], c); cur += c; i++; }
      }
      return cur == n;
    }
    case 0: {                                             /* HUFFMAN_NORMAL: code table (bit stuffed like codec version >= 3) + MSB-first bit stream */
      uint16_t len[256]; u32 code[256]; int sz = 0;
      size_t tb = huff_read_table(p + 1, size - 1, len, code, &sz, 5);
      if (!tb) return 0;
      const u8* q = p + 1 + tb; size_t avail = size - 1 - tb;
      trie* t = (trie*)malloc(sizeof(trie));
      if (!trie_build(t, len, code, sz)) { free(t); return 0; }
      u64 pos = 0, nBits = (u64)(avail / 4) * 32;
      int ok = 1;
      for (size_t m = 0; m < n && ok; m++) {
        int cur = 0;
        while (t->n[cur].sym < 0) {
          if (pos >= nBits) { ok = 0; break; }
          int nx = t->n[cur].kid[msb_get(q, pos++)];
          if (nx < 0) { ok = 0; break; }
          cur = nx;
        }
        if (ok) out[m] = (u8)t->n[cur].sym;
      }
      free(t);
      /* like the 8-bit Huffman path (Lerc2.cpp:2583-2587): the data words plus the read-ahead word must be there.  The reference's
       * FPL reader does not check this (it would read past the plane on corrupted input); the oracle and the product do. */
      if (ok && avail < 4 * (((pos & 31) ? 1 : 0) + (size_t)(pos >> 5) + 1)) ok = 0;
      return ok;
    }
    default: return 0;
  }
}

static u32 fpl_add32(u32 a, u32 b) { return ((a + b) & 0x007FFFFFu) | (((((a >> 23) & 0x1FF) + ((b >> 23) & 0x1FF)) & 0x1FF) << 23); }
static u64 fpl_add64(u64 a, u64 b) { return ((a + b) & 0x000FFFFFFFFFFFFFull) | (((((a >> 52) & 0xFFF) + ((b >> 52) & 0xFFF)) & 0xFFF) << 52); }

/* One slice of cols x rows units.  Returns the number of bytes consumed, 0 on failure. */
static size_t fpl_decode_slice(const u8* p, size_t avail, void* data, int isDouble, size_t cols, size_t rows) {
  const u8* p0 = p;
  const size_t unit = isDouble ? 8 : 4, n = cols * rows;
  if (avail < 1 || n == 0) return 0;
  const int pred = *p++; avail--;
  if (pred > 2) return 0;
  u8* planes = (u8*)malloc(n * unit);
  int ok = 1;
  u8* asm_ = (u8*)data;
  for (size_t b = 0; b < unit && ok; b++) {
    if (avail < 6) { ok = 0; break; }
    u8 idx = p[0], level = p[1]; u32 csize; memcpy(&csize, p + 2, 4);
    p += 6; avail -= 6;
    if (idx >= unit || level > 5 || avail < csize) { ok = 0; break; }
    u8* pl = planes + b * n;
    if (!fpl_plane_decode(p, csize, pl, n)) { ok = 0; break; }
    p += csize; avail -= csize;
    for (int l = level; l > 0; l--) for (size_t i = (size_t)l; i < n; i++) pl[i] = (u8)(pl[i] + pl[i - 1]);     /* restoreSequence */
    for (size_t i = 0; i < n; i++) asm_[i * unit + idx] = pl[i];
  }
  free(planes);
  if (!ok) return 0;
  const int delta = pred == 1 ? 1 : (pred == 2 ? 2 : 0);
  if (!isDouble) {
    u32* d = (u32*)data;
    if (pred == 2) for (size_t c = 0; c < cols; c++) for (size_t r = 1; r < rows; r++) d[r * cols + c] = fpl_add32(d[r * cols + c], d[(r - 1) * cols + c]);
    if (delta > 0) for (size_t r = 0; r < rows; r++) for (size_t i = 1; i < cols; i++) d[r * cols + i] = fpl_add32(d[r * cols + i], d[r * cols + i - 1]);
    for (size_t i = 0; i < n; i++) { u32 a = d[i]; d[i] = (a & 0x007FFFFFu) | (((a >> 24) & 0xFF) << 23) | (((a >> 23) & 1u) << 31); }   /* undo_moveBits2Front */
  } else {
    u64* d = (u64*)data;
    if (pred == 2) for (size_t c = 0; c < cols; c++) for (size_t r = 1; r < rows; r++) d[r * cols + c] = fpl_add64(d[r * cols + c], d[(r - 1) * cols + c]);
    if (delta > 0) for (size_t r = 0; r < rows; r++) for (size_t i = 1; i < cols; i++) d[r * cols + i] = fpl_add64(d[r * cols + i], d[r * cols + i - 1]);
  }
  return (size_t)(p - p0);
}

/* fpl_Lerc2Ext.cpp:725-736: nDepth > 1 is one slice of nDepth columns and nCols * nRows rows */
static size_t fpl_decode(const u8* p, size_t avail, void* data, int isDouble, int nCols, int nRows, int nDepth) {
  if (nDepth == 1) return fpl_decode_slice(p, avail, data, isDouble, (size_t)nCols, (size_t)nRows);
  return fpl_decode_slice(p, avail, data, isDouble, (size_t)nDepth, (size_t)nCols * (size_t)nRows);
}

/* ------------------------------------------------------------------------------------------- */
/* FPL, ENCODE side.  fpl_Lerc2Ext.cpp:62-101 (test blocks), :103-131 (byte derivatives), :170-232 (testBlocksSize), :238-330
 * (getBestLevel2), :341-395 (selectInitialLinearOrCrossDelta), :397-436 (compressedLength / EncodeHuffmanFlt), :438-608
 * (ComputeHuffmanCodesFlt[Slice]); fpl_Compression.cpp:53-112 (compress_buffer / getEntropySize); fpl_EsriHuffman.cpp:82-259
 * (PackBits), :262-452 (EncodeHuffman); fpl_UnitTypes.cpp:39-51, :83-97, :119-136, :302-357, :436-517 (float transform, split
 * subtraction, row / cross derivatives); fpl_Predictor.cpp:30-76.
 *
 * One deviation, on purpose: the reference leaves the read-ahead word behind every Huffman-coded byte plane uninitialised (it
 * is malloc'ed and never written, fpl_EsriHuffman.cpp:403-448, and copied into the blob with the plane).  This restatement
 * writes zeros there, like the 8-bit Huffman path does; tests compare reference-made blobs with those 4 bytes per plane (and the
 * checksum they feed) set aside. */

enum { FPL_PRIME = 7, FPL_MAX_DELTA = 5, FPL_SAMPLE = 8 * 1024 };

/* Test switch: 0 makes the encoder skip the lossless float codec (raw tiling / one sweep is written instead, as a build without
 * fpl_*.cpp would), for checking a product that does not write FPL yet.  Returns the previous setting.  Default: on = the reference. */
static int g_fplEncoder = 1;
int lo_fpl_encoder(int on) { int was = g_fplEncoder; g_fplEncoder = on != 0; return was; }

typedef struct { u8 pred; int nPlanes; u8 level[8]; u32 size[8]; u8* buf[8]; } fpl_plan;

static void fpl_plan_free(fpl_plan* pl) { for (int i = 0; i < 8; i++) { free(pl->buf[i]); pl->buf[i] = NULL; } pl->nPlanes = 0; }

static u32 fpl_sub32(u32 a, u32 b) { return ((a - b) & 0x007FFFFFu) | (((((a >> 23) & 0x1FF) - ((b >> 23) & 0x1FF)) & 0x1FF) << 23); }
static u64 fpl_sub64(u64 a, u64 b) {
  return (((a & 0x000FFFFFFFFFFFFFull) - (b & 0x000FFFFFFFFFFFFFull)) & 0x000FFFFFFFFFFFFFull) | (((((a >> 52) & 0xFFF) - ((b >> 52) & 0xFFF)) & 0xFFF) << 52);
}
static u32 fpl_bits_to_front(u32 a) { return (a & 0x007FFFFFu) | (((a >> 23) & 0xFF) << 24) | ((a >> 31) << 23); }

/* fpl_Compression.cpp:85-112: entropy of every 7th byte, in whole bytes */
static long fpl_entropy_size(const u8* p, size_t size) {
  unsigned long table[256]; memset(table, 0, sizeof table);
  int total = 0;
  for (size_t i = 0; i < size; i += FPL_PRIME) { table[p[i]]++; total++; }
  double bitsSum = 0;
  for (int i = 0; i < 256; i++) {
    if (!table[i]) continue;
    double pr = (double)total / table[i];
    double bits = log2(pr);
    bitsSum += bits * table[i];
  }
  return (long)((bitsSum + 7) / 8);
}

/* row derivative (level 1 inside every row) and column derivative; both run backwards, so every element sees its
 * neighbour's old value */
static void fpl_rows_derivative(void* data, int isDouble, size_t cols, size_t rows) {
  for (size_t r = 0; r < rows; r++)
    for (size_t i = cols - 1; i >= 1; i--) {
      if (isDouble) { u64* d = (u64*)data + r * cols; d[i] = fpl_sub64(d[i], d[i - 1]); }
      else { u32* d = (u32*)data + r * cols; d[i] = fpl_sub32(d[i], d[i - 1]); }
    }
}
static void fpl_cols_derivative(void* data, int isDouble, size_t cols, size_t rows) {
  for (size_t c = 0; c < cols; c++)
    for (size_t r = rows - 1; r >= 1; r--) {
      if (isDouble) { u64* d = (u64*)data; d[r * cols + c] = fpl_sub64(d[r * cols + c], d[(r - 1) * cols + c]); }
      else { u32* d = (u32*)data; d[r * cols + c] = fpl_sub32(d[r * cols + c], d[(r - 1) * cols + c]); }
    }
}

typedef struct { long top, height; } fpl_block;

/* fpl_Lerc2Ext.cpp:62-101.  Returns the number of blocks (<= count). */
static int fpl_test_blocks(int width, int height, fpl_block** out) {
  size_t size = (size_t)width * (size_t)height;
  double t = round((double)size / FPL_SAMPLE);
  int count = (int)round(sqrt(t + 1));
  int bh = FPL_SAMPLE / width;
  if (bh < 4) bh = 4;
  while (count * bh > height && count > 1) count--;
  float topMargin = (float)((height - count * bh) / (2.0 * count));
  float delta = 2.0f * topMargin + bh;
  fpl_block* b = (fpl_block*)malloc((size_t)count * sizeof(fpl_block));
  int n = 0;
  for (int i = 0; i < count; i++) {
    fpl_block tb; tb.top = (long)(topMargin + delta * i); tb.height = bh;
    if (tb.top < 0) tb.top = 0;
    if (tb.top + tb.height > height) tb.height = height - tb.top;
    if (tb.height > 0) b[n++] = tb;
  }
  *out = b;
  return n;
}

/* fpl_Lerc2Ext.cpp:170-232 with test_first_byte_delta = true */
static size_t fpl_test_blocks_size(const fpl_block* blk, int nBlk, int unit, const u8* data, long width) {
  size_t ret = 0;
  for (int b = 0; b < nBlk; b++) {
    size_t start = (size_t)unit * (size_t)blk[b].top * (size_t)width, length = (size_t)blk[b].height * (size_t)width;
    u8* plane = (u8*)malloc(length);
    for (int byte = 0; byte < unit; byte++) {
      for (size_t i = 0; i < length; i++) plane[i] = data[start + (size_t)byte + i * (size_t)unit];
      size_t e1 = (size_t)fpl_entropy_size(plane, length);
      int off = FPL_PRIME * (((int)length - 1) / FPL_PRIME);                    /* setDerivativePrime :103-116 */
      for (; off >= 1; off -= FPL_PRIME) plane[off] = (u8)(plane[off] - plane[off - 1]);
      size_t e2 = (size_t)fpl_entropy_size(plane, length);
      ret += e1 < e2 ? e1 : e2;
    }
    free(plane);
  }
  return ret;
}

/* fpl_Lerc2Ext.cpp:238-330 */
static int fpl_best_level(const u8* p, size_t size, int maxDelta) {
  if (maxDelta == 0) return 0;
  const unsigned target = FPL_SAMPLE;
  double t = round((double)size / target);
  int count = (int)round(sqrt(t + 1));
  while ((size_t)((unsigned)count * target) > size && count > 0) count--;
  if (count == 0) return 0;                       /* no snippet: every level estimates 0 bytes, level 0 stays */
  float topMargin = (float)(((unsigned)(int)size - (unsigned)count * target) / (2.0 * count));
  float delta = 2.0f * topMargin + target;
  long* start = (long*)malloc((size_t)count * sizeof(long)); int* len = (int*)malloc((size_t)count * sizeof(int));
  int n = 0;
  for (int i = 0; i < count; i++) {
    long st = (long)(topMargin + delta * i); int ln = (int)target;
    if (st < 0) st = 0;
    if (st + ln > (int)size) ln = (int)size - (int)st;
    if (ln > 0) { start[n] = st; len[n] = ln; n++; }
  }
  u8* copy = (u8*)malloc(size); memcpy(copy, p, size);
  size_t best = 0; int ret = 0;
  for (int l = 0; l <= maxDelta; l++) {
    if (l > 0)
      for (int s = 0; s < n; s++)
        for (int i = (int)start[s] + len[s] - 1; i >= (int)start[s] + l; i--) copy[i] = (u8)(copy[i] - copy[i - 1]);
    size_t comp = 0;
    for (int s = 0; s < n; s++) comp += (size_t)fpl_entropy_size(copy + start[s], (size_t)len[s]);
    if (comp < best || l == 0) { best = comp; ret = l; } else break;
  }
  free(copy); free(start); free(len);
  return ret;
}

/* fpl_EsriHuffman.cpp:82-166 / :169-259.  out == NULL: size only. */
static long fpl_packbits_encode(const u8* ptr, size_t size, u8* out) {
  long curr = 0, litPos = -1; int lit = 0;
  for (size_t i = 0; i <= size;) {
    int b = (i == size) ? -1 : ptr[i];
    int rep = 0;
    while (i < size - 1 && b == ptr[i + 1] && rep < 128) { i++; rep++; }
    i++;
    if (rep == 0 && b >= 0) {
      if (litPos < 0) { litPos = curr; curr++; }
      if (out) out[curr] = (u8)b;
      curr++; lit++;
      if (lit == 128) { if (out) out[litPos] = (u8)(lit - 1); lit = 0; litPos = -1; }
    } else {
      if (lit > 0) { if (out) out[litPos] = (u8)(lit - 1); litPos = -1; lit = 0; }
      if (rep > 0) { if (out) { out[curr] = (u8)(127 + rep); out[curr + 1] = (u8)b; } curr += 2; }
    }
  }
  return curr;
}

/* fpl_EsriHuffman.cpp:316-452 with use_rle.  Returns a malloc'ed buffer and its size; 0 = failure (:340-343). */
static size_t fpl_plane_encode(const u8* in, size_t n, u8** outBuf) {
  int histo[256]; memset(histo, 0, sizeof histo);
  for (size_t i = 0; i < n; i++) histo[in[i]]++;
  int distinct = 0;
  for (int i = 0; i < 256; i++) distinct += histo[i] > 0;
  if (distinct < 2) {                                                           /* one value: flag, value, count */
    u8* o = (u8*)calloc(6, 1); u32 len = (u32)n;
    o[0] = 1; o[1] = in[0]; memcpy(o + 2, &len, 4);
    *outBuf = o; return 6;
  }
  uint16_t hl[256]; u32 hc[256]; int numBytes = 0;
  if (!lo_huffman_lengths(histo, 256, hl, hc) || !huff_total_bytes(histo, hl, 256, &numBytes) || numBytes <= 0) return 0;
  long pb = fpl_packbits_encode(in, n, NULL);
  if (pb > 0 && pb < numBytes && pb < (long)n) {
    u8* o = (u8*)malloc((size_t)pb + 1);
    o[0] = 3; fpl_packbits_encode(in, n, o + 1);
    *outBuf = o; return (size_t)pb + 1;
  }
  if (numBytes >= (int)n) {
    u8* o = (u8*)malloc(n + 1);
    o[0] = 2; memcpy(o + 1, in, n);
    *outBuf = o; return n + 1;
  }
  u8* o = (u8*)calloc((size_t)numBytes + 1, 1);
  o[0] = 0;
  size_t tb = huff_write_table(o + 1, hl, hc, 256, 5);
  if (!tb) { free(o); return 0; }
  msb_writer w = {o + 1 + tb, 0};
  for (size_t i = 0; i < n; i++) msb_put(&w, hc[in[i]], hl[in[i]]);
  size_t total = 1 + tb + 4 * ((size_t)(w.bitPos >> 5) + ((w.bitPos & 31) ? 1 : 0) + 1);
  *outBuf = o; return total;
}

/* fpl_Lerc2Ext.cpp:455-608 */
static int fpl_plan_slice(const void* input, int isDouble, size_t cols, size_t rows, fpl_plan* plan) {
  const int unit = isDouble ? 8 : 4;
  const size_t n = cols * rows;
  fpl_plan_free(plan);
  u8* values = (u8*)malloc(n * (size_t)unit); memcpy(values, input, n * (size_t)unit);
  if (!isDouble) { u32* v = (u32*)values; for (size_t i = 0; i < n; i++) v[i] = fpl_bits_to_front(v[i]); }
  u8* copy = (u8*)malloc(n * (size_t)unit); memcpy(copy, values, n * (size_t)unit);
  fpl_block* blk = NULL; int nBlk = fpl_test_blocks((int)cols, (int)rows, &blk);
  size_t stats[3];
  stats[0] = fpl_test_blocks_size(blk, nBlk, unit, copy, (long)cols);
  fpl_rows_derivative(copy, isDouble, cols, rows);
  stats[1] = fpl_test_blocks_size(blk, nBlk, unit, copy, (long)cols);
  fpl_cols_derivative(copy, isDouble, cols, rows);
  stats[2] = fpl_test_blocks_size(blk, nBlk, unit, copy, (long)cols);
  free(copy); free(blk);
  int pred = 0;
  if (stats[1] < stats[pred]) pred = 1;
  if (stats[2] < stats[pred]) pred = 2;
  if (pred >= 1) fpl_rows_derivative(values, isDouble, cols, rows);
  if (pred == 2) fpl_cols_derivative(values, isDouble, cols, rows);
  const int maxDelta = FPL_MAX_DELTA - pred;
  u8* plane = (u8*)malloc(n);
  int ok = 1;
  plan->pred = (u8)pred;
  for (int byte = 0; byte < unit && ok; byte++) {
    for (size_t i = 0; i < n; i++) plane[i] = values[i * (size_t)unit + (size_t)byte];
    int level = fpl_best_level(plane, n, maxDelta);
    for (int l = 1; l <= level; l++) for (int i = (int)n - 1; i >= l; i--) plane[i] = (u8)(plane[i] - plane[i - 1]);   /* :118-131 */
    size_t sz = fpl_plane_encode(plane, n, &plan->buf[byte]);
    if (!sz) { ok = 0; break; }
    plan->level[byte] = (u8)level; plan->size[byte] = (u32)sz; plan->nPlanes = byte + 1;
  }
  free(plane); free(values);
  if (!ok) fpl_plan_free(plan);
  return ok;
}

static int fpl_plan_make(const void* input, int isDouble, int nCols, int nRows, int nDepth, fpl_plan* plan) {   /* :438-453 */
  if (nDepth == 1) return fpl_plan_slice(input, isDouble, (size_t)nCols, (size_t)nRows, plan);
  return fpl_plan_slice(input, isDouble, (size_t)nDepth, (size_t)nCols * (size_t)nRows, plan);
}
static int fpl_plan_bytes(const fpl_plan* pl) { int64_t r = 1; for (int i = 0; i < pl->nPlanes; i++) r += (int64_t)pl->size[i] + 6; return r > INT_MAX ? -1 : (int)r; }   /* :397-408 */
static u8* fpl_plan_write(const fpl_plan* pl, u8* p) {                                                              /* :410-436 */
  *p++ = pl->pred;
  for (int i = 0; i < pl->nPlanes; i++) {
    *p++ = (u8)i; *p++ = pl->level[i];
    memcpy(p, &pl->size[i], 4); p += 4;
    memcpy(p, pl->buf[i], pl->size[i]); p += pl->size[i];
  }
  return p;
}

/* ------------------------------------------------------------------------------------------- */
/* shared helpers of the typed code                                                               */

static int mask_bit(const u8* bits, int64_t k) { return (bits[k >> 3] & (0x80 >> (k & 7))) != 0; }

/* BitMask.cpp:100-119 */
static int64_t mask_count(const u8* bits, int64_t nPix) {
  int64_t c = 0;
  for (int64_t k = 0; k < nPix; k++) c += mask_bit(bits, k);
  return c;
}

/* Lerc2.h:457-542: smallest type that stores the block offset exactly; returns the 2-bit code. */
static int fits_int(double z, double lo, double hi) { return z >= lo && z <= hi && z == floor(z); }
static int reduce_offset_type(double z, int dt, int* dtUsed) {
  int tc = 0;
  switch (dt) {
    case LO_SHORT:  tc = fits_int(z, -128, 127) ? 2 : (fits_int(z, 0, 255) ? 1 : 0); *dtUsed = dt - tc; break;
    case LO_USHORT: tc = fits_int(z, 0, 255) ? 1 : 0; *dtUsed = dt - 2 * tc; break;
    case LO_INT:    tc = fits_int(z, 0, 255) ? 3 : (fits_int(z, -32768, 32767) ? 2 : (fits_int(z, 0, 65535) ? 1 : 0)); *dtUsed = dt - tc; break;
    case LO_UINT:   tc = fits_int(z, 0, 255) ? 2 : (fits_int(z, 0, 65535) ? 1 : 0); *dtUsed = dt - 2 * tc; break;
    case LO_FLOAT:  tc = fits_int(z, 0, 255) ? 2 : (fits_int(z, -32768, 32767) ? 1 : 0); *dtUsed = tc == 0 ? dt : (tc == 1 ? LO_SHORT : LO_BYTE); break;
    case LO_DOUBLE:
      tc = fits_int(z, -32768, 32767) ? 3 : (fits_int(z, (double)INT_MIN, (double)INT_MAX) ? 2 : ((z >= -FLT_MAX && z <= FLT_MAX && (double)(float)z == z) ? 1 : 0));
      *dtUsed = tc == 0 ? dt : dt - 2 * tc + 1; break;
    default: *dtUsed = dt; break;
  }
  return tc;
}

static int offset_type_from_code(int dt, int tc) {      /* Lerc2.h:528-542 */
  int r;
  switch (dt) {
    case LO_SHORT: case LO_INT: r = dt - tc; break;
    case LO_USHORT: case LO_UINT: r = dt - 2 * tc; break;
    case LO_FLOAT: r = tc == 0 ? dt : (tc == 1 ? LO_SHORT : LO_BYTE); break;
    case LO_DOUBLE: r = tc == 0 ? dt : dt - 2 * tc + 1; break;
    default: r = dt; break;
  }
  return (r >= LO_CHAR && r <= LO_DOUBLE) ? r : LO_UNDEFINED;
}

static u8* write_offset(u8* p, double z, int dtUsed) {   /* Lerc2.h:546-613 */
  switch (dtUsed) {
    case LO_CHAR:   { int8_t v = (int8_t)z; memcpy(p, &v, 1); return p + 1; }
    case LO_BYTE:   { u8 v = (u8)z; memcpy(p, &v, 1); return p + 1; }
    case LO_SHORT:  { int16_t v = (int16_t)z; memcpy(p, &v, 2); return p + 2; }
    case LO_USHORT: { uint16_t v = (uint16_t)z; memcpy(p, &v, 2); return p + 2; }
    case LO_INT:    { int32_t v = (int32_t)z; memcpy(p, &v, 4); return p + 4; }
    case LO_UINT:   { u32 v = (u32)z; memcpy(p, &v, 4); return p + 4; }
    case LO_FLOAT:  { float v = (float)z; memcpy(p, &v, 4); return p + 4; }
    default:        { memcpy(p, &z, 8); return p + 8; }
  }
}

static double read_offset(const u8* p, int dtUsed) {     /* Lerc2.h:617-681 */
  switch (dtUsed) {
    case LO_CHAR:   { int8_t v; memcpy(&v, p, 1); return v; }
    case LO_BYTE:   { return *p; }
    case LO_SHORT:  { int16_t v; memcpy(&v, p, 2); return v; }
    case LO_USHORT: { uint16_t v; memcpy(&v, p, 2); return v; }
    case LO_INT:    { int32_t v; memcpy(&v, p, 4); return v; }
    case LO_UINT:   { u32 v; memcpy(&v, p, 4); return v; }
    case LO_FLOAT:  { float v; memcpy(&v, p, 4); return v; }
    default:        { double v; memcpy(&v, p, 8); return v; }
  }
}

static u32 max_val_to_quantize(int dt) { return dt <= LO_USHORT ? (1u << 15) - 1 : (1u << 30) - 1; }  /* Lerc2.h:685-703 */

/* per-band encoder/decoder state shared between the typed functions (the role of class Lerc2) */
typedef struct {
  lo_hdr hd;
  u8* bits;              /* bit mask, (nPix+7)/8 bytes, MSB first */
  int haveBits;          /* decoder: a mask from a previous band may be reused (Lerc2.cpp:1002) */
  double* zMinVec; double* zMaxVec;   /* per depth */
  int minMaxSet;         /* Lerc.cpp:749-751 */
  int encodeMask, oneSweep, imageMode;
  u32 maxQ;
  uint16_t hLen[256]; u32 hCode[256]; int haveHuff;
  fpl_plan fpl;          /* lossless float codec: the coded byte planes wait here between sizing and writing */
} band_state;

enum { IEM_TILING = 0, IEM_DELTA_HUFFMAN = 1, IEM_HUFFMAN = 2, IEM_DELTA_DELTA_HUFFMAN = 3 };

static int try_huffman_int(const lo_hdr* h) { return h->version >= 2 && (h->dt == LO_BYTE || h->dt == LO_CHAR) && h->maxZErr == 0.5; }
static int try_huffman_flt(const lo_hdr* h) { return h->version >= 6 && (h->dt == LO_FLOAT || h->dt == LO_DOUBLE) && h->maxZErr == 0; }

/* Lerc2.cpp:1012-1030 */
static int finish_blob(u8* blob, const u8* end, const lo_hdr* h) {
  if ((size_t)(end - blob) != (size_t)h->blobSize) return 0;
  if (h->version >= 3) {
    u32 cs = lo_fletcher32(blob + 14, h->blobSize - 14);
    memcpy(blob + 10, &cs, 4);
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* typed instantiations                                                                           */

static int read_band_header(const u8* p, size_t avail, lo_hdr* h, int* hasMask);

#define T int8_t
#define TN(x) x##_i8
#define T_CODE LO_CHAR
#define T_IS_FLT 0
#include "lerc_oracle_typed.inc"
#undef T
#undef TN
#undef T_CODE
#undef T_IS_FLT

#define T uint8_t
#define TN(x) x##_u8
#define T_CODE LO_BYTE
#define T_IS_FLT 0
#include "lerc_oracle_typed.inc"
#undef T
#undef TN
#undef T_CODE
#undef T_IS_FLT

#define T int16_t
#define TN(x) x##_i16
#define T_CODE LO_SHORT
#define T_IS_FLT 0
#include "lerc_oracle_typed.inc"
#undef T
#undef TN
#undef T_CODE
#undef T_IS_FLT

#define T uint16_t
#define TN(x) x##_u16
#define T_CODE LO_USHORT
#define T_IS_FLT 0
#include "lerc_oracle_typed.inc"
#undef T
#undef TN
#undef T_CODE
#undef T_IS_FLT

#define T int32_t
#define TN(x) x##_i32
#define T_CODE LO_INT
#define T_IS_FLT 0
#include "lerc_oracle_typed.inc"
#undef T
#undef TN
#undef T_CODE
#undef T_IS_FLT

#define T uint32_t
#define TN(x) x##_u32
#define T_CODE LO_UINT
#define T_IS_FLT 0
#include "lerc_oracle_typed.inc"
#undef T
#undef TN
#undef T_CODE
#undef T_IS_FLT

#define T float
#define TN(x) x##_f32
#define T_CODE LO_FLOAT
#define T_IS_FLT 1
#include "lerc_oracle_typed.inc"
#undef T
#undef TN
#undef T_CODE
#undef T_IS_FLT

#define T double
#define TN(x) x##_f64
#define T_CODE LO_DOUBLE
#define T_IS_FLT 1
#include "lerc_oracle_typed.inc"
#undef T
#undef TN
#undef T_CODE
#undef T_IS_FLT

/* ------------------------------------------------------------------------------------------- */
/* API                                                                  Lerc_c_api_impl.cpp:33-305 */

typedef unsigned (*enc_fn)(const void*, int, int, int, int, int, int, const u8*, double, u8*, unsigned, unsigned*, unsigned*, const u8*, const double*);
typedef unsigned (*dec_fn)(const u8*, unsigned, int, u8*, int, int, int, int, void*, u8*, double*);
static const enc_fn kEnc[8] = {encode_bands_i8, encode_bands_u8, encode_bands_i16, encode_bands_u16,
                               encode_bands_i32, encode_bands_u32, encode_bands_f32, encode_bands_f64};
static const dec_fn kDec[8] = {decode_bands_i8, decode_bands_u8, decode_bands_i16, decode_bands_u16,
                               decode_bands_i32, decode_bands_u32, decode_bands_f32, decode_bands_f64};

static int dims_ok(int nDepth, int nCols, int nRows, int elemSize) {     /* Lerc.cpp:1622-1639 */
  if (nDepth <= 0 || nCols <= 0 || nRows <= 0) return 0;
  u64 nPix = (u64)nRows * (u64)nCols, lim = (u64)INT_MAX, b = (u64)elemSize;
  return !(nPix > lim || b * (u64)nDepth > lim || b * (u64)nDepth * nPix > lim);
}

static unsigned encode_common4(const void* data, int version, unsigned dt, int nDepth, int nCols, int nRows, int nBands,
                               int nMasks, const u8* validBytes, double maxZErr, u8* out, unsigned outSize,
                               unsigned* nWritten, unsigned* nNeeded, int sizeOnly, const u8* usesNoData, const double* noDataValues);
static unsigned encode_common(const void* data, int version, unsigned dt, int nDepth, int nCols, int nRows, int nBands,
                              int nMasks, const u8* validBytes, double maxZErr, u8* out, unsigned outSize,
                              unsigned* nWritten, unsigned* nNeeded, int sizeOnly) {
  return encode_common4(data, version, dt, nDepth, nCols, nRows, nBands, nMasks, validBytes, maxZErr, out, outSize, nWritten, nNeeded, sizeOnly, NULL, NULL);
}
static unsigned encode_common4(const void* data, int version, unsigned dt, int nDepth, int nCols, int nRows, int nBands,
                               int nMasks, const u8* validBytes, double maxZErr, u8* out, unsigned outSize,
                               unsigned* nWritten, unsigned* nNeeded, int sizeOnly, const u8* usesNoData, const double* noDataValues) {
  if (!data || dt >= LO_UNDEFINED || nDepth <= 0 || nCols <= 0 || nRows <= 0 || nBands <= 0 || maxZErr < 0) return LO_WRONG_PARAM;
  if (!sizeOnly && (!out || !outSize)) return LO_WRONG_PARAM;
  if (!(nMasks == 0 || nMasks == 1 || nMasks == nBands) || (nMasks > 0 && !validBytes)) return LO_WRONG_PARAM;
  if (version > 6 || (version >= 0 && version < 2)) return LO_WRONG_PARAM;   /* Lerc2::SetEncoderToOldVersion, Lerc2.cpp:52-63 */
  if (!dims_ok(nDepth, nCols, nRows, kTypeSize[dt])) return LO_DIMS_TOO_LARGE;
  if (!sizeOnly) memset(out, 0, outSize);                          /* Lerc.cpp:374 */
  int anyNoData = 0;
  if (usesNoData) for (int i = 0; i < nBands; i++) if (usesNoData[i]) anyNoData = 1;
  if (anyNoData && (!noDataValues || (version >= 0 && version <= 5))) return LO_WRONG_PARAM;     /* Lerc.cpp:649-652, :378-383 */
  return kEnc[dt](data, version < 0 ? 6 : version, nDepth, nCols, nRows, nBands, nMasks, validBytes, maxZErr, sizeOnly ? NULL : out, outSize, nWritten, nNeeded,
                  anyNoData ? usesNoData : NULL, noDataValues);
}

unsigned lo_computeCompressedSizeForVersion(const void* data, int version, unsigned dt, int nDepth, int nCols, int nRows,
                                            int nBands, int nMasks, const unsigned char* validBytes, double maxZErr, unsigned* numBytes) {
  if (!numBytes) return LO_WRONG_PARAM;
  *numBytes = 0;
  unsigned w = 0;
  return encode_common(data, version, dt, nDepth, nCols, nRows, nBands, nMasks, validBytes, maxZErr, NULL, 0, &w, numBytes, 1);
}
unsigned lo_computeCompressedSize(const void* data, unsigned dt, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                                  const unsigned char* validBytes, double maxZErr, unsigned* numBytes) {
  return lo_computeCompressedSizeForVersion(data, -1, dt, nDepth, nCols, nRows, nBands, nMasks, validBytes, maxZErr, numBytes);
}
unsigned lo_encodeForVersion(const void* data, int version, unsigned dt, int nDepth, int nCols, int nRows, int nBands,
                             int nMasks, const unsigned char* validBytes, double maxZErr, unsigned char* out, unsigned outSize,
                             unsigned* nWritten) {
  if (!nWritten) return LO_WRONG_PARAM;
  *nWritten = 0;
  unsigned need = 0;
  return encode_common(data, version, dt, nDepth, nCols, nRows, nBands, nMasks, validBytes, maxZErr, out, outSize, nWritten, &need, 0);
}
unsigned lo_encode(const void* data, unsigned dt, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                   const unsigned char* validBytes, double maxZErr, unsigned char* out, unsigned outSize, unsigned* nWritten) {
  return lo_encodeForVersion(data, -1, dt, nDepth, nCols, nRows, nBands, nMasks, validBytes, maxZErr, out, outSize, nWritten);
}

/* the multi-band header walk of Lerc::GetLercInfo                                Lerc.cpp:92-182 */
typedef struct {
  int version, nDepth, nCols, nRows, numValid, nBands, nMasks, dt, nUsesNoData;
  unsigned blobSize;
  double zMin, zMax, maxZErr;
} lo_info;

static int read_band_header(const u8* p, size_t avail, lo_hdr* h, int* hasMask) {   /* Lerc2.cpp:495-510 */
  int n = hdr_read(p, avail, h);
  if (!n || avail - (size_t)n < 4) return 0;
  int nm; memcpy(&nm, p + n, 4);
  if (nm < 0) return 0;
  *hasMask = nm > 0;
  return 1;
}

/* ranges of one band into mins/maxs[iBand*nDepth ...]                Lerc.cpp:1014-1042, Lerc2.cpp:514-573 */
static unsigned band_ranges(const u8* p, size_t avail, int iBand, const lo_hdr* h, double* mins, double* maxs, size_t nElem) {
  int nDepth = h->nDepth;
  if (nElem < ((size_t)iBand + 1) * (size_t)nDepth) return LO_BUFFER_TOO_SMALL;
  if (nDepth == 1) { mins[iBand] = h->zMin; maxs[iBand] = h->zMax; return LO_OK; }
  if (h->passNoData) return LO_HAS_NODATA;
  if (h->version < 4) return LO_FAILED;
  double* mn = mins + (size_t)iBand * nDepth; double* mx = maxs + (size_t)iBand * nDepth;
  if (h->numValid == 0) { for (int i = 0; i < nDepth; i++) mn[i] = mx[i] = 0; return LO_OK; }
  if (h->zMin == h->zMax) { for (int i = 0; i < nDepth; i++) mn[i] = mx[i] = h->zMin; return LO_OK; }
  size_t off = (size_t)hdr_bytes(h->version);
  if (avail < off + 4) return LO_FAILED;
  int nm; memcpy(&nm, p + off, 4);
  off += 4;
  if (nm < 0 || avail < off + (size_t)nm) return LO_FAILED;
  off += (size_t)nm;
  size_t ts = (size_t)kTypeSize[h->dt], len = ts * (size_t)nDepth;
  if (avail < off + 2 * len) return LO_FAILED;
  for (int i = 0; i < nDepth; i++) {
    mn[i] = read_offset(p + off + ts * (size_t)i, h->dt);
    mx[i] = read_offset(p + off + len + ts * (size_t)i, h->dt);
  }
  return LO_OK;
}

static unsigned blob_info(const u8* blob, unsigned blobSize, lo_info* li, double* mins, double* maxs, size_t nElem) {
  memset(li, 0, sizeof *li);
  lo_hdr h; int hasMask = 0, nMasks = 0;
  if (!read_band_header(blob, blobSize, &h, &hasMask)) return LO_FAILED;     /* (no Lerc1 in the oracle) */
  li->version = h.version; li->nDepth = h.nDepth; li->nCols = h.nCols; li->nRows = h.nRows;
  li->numValid = h.numValid; li->blobSize = (unsigned)h.blobSize; li->dt = h.dt;
  li->zMin = h.zMin; li->zMax = h.zMax; li->maxZErr = h.maxZErr; li->nUsesNoData = h.passNoData ? 1 : 0;
  int tryNext = h.version <= 5 || h.nBlobsMore > 0;
  if (hasMask || h.numValid == 0) nMasks = 1;
  if (mins && maxs) { unsigned e = band_ranges(blob, blobSize, 0, &h, mins, maxs, nElem); if (e) return e; }
  li->nBands = 1;
  if (li->blobSize > blobSize) return LO_FAILED;
  lo_hdr g;
  while (tryNext && read_band_header(blob + li->blobSize, blobSize - li->blobSize, &g, &hasMask)) {
    if (g.nDepth != li->nDepth || g.nCols != li->nCols || g.nRows != li->nRows || g.dt != li->dt) return LO_FAILED;
    tryNext = g.version <= 5 || g.nBlobsMore > 0;
    if (g.passNoData) li->nUsesNoData++;
    if (hasMask || g.numValid != li->numValid) nMasks = 2;
    if ((u64)li->blobSize + (u64)g.blobSize > (u64)UINT_MAX) return LO_FAILED;
    if ((u64)li->blobSize + (u64)g.blobSize > (u64)blobSize) return LO_FAILED;
    if (g.zMin < li->zMin) li->zMin = g.zMin;
    if (g.zMax > li->zMax) li->zMax = g.zMax;
    if (g.maxZErr > li->maxZErr) li->maxZErr = g.maxZErr;
    if (mins && maxs) {
      unsigned e = band_ranges(blob + li->blobSize, blobSize - li->blobSize, li->nBands, &g, mins, maxs, nElem);
      if (e) return e;
    }
    li->blobSize += (unsigned)g.blobSize;
    li->nBands++;
  }
  li->nMasks = nMasks > 1 ? li->nBands : nMasks;
  if (li->nUsesNoData > 0) li->nUsesNoData = li->nBands;
  return LO_OK;
}

unsigned lo_getBlobInfo(const unsigned char* blob, unsigned blobSize, unsigned* infoArray, double* dataRangeArray,
                        int infoArraySize, int dataRangeArraySize) {
  if (!blob || !blobSize || (!infoArray && !dataRangeArray) || (infoArraySize <= 0 && dataRangeArraySize <= 0)) return LO_WRONG_PARAM;
  lo_info li;
  unsigned e = blob_info(blob, blobSize, &li, NULL, NULL, 0);
  if (e) return e;
  if (infoArray) {
    unsigned v[11] = {(unsigned)li.version, (unsigned)li.dt, (unsigned)li.nDepth, (unsigned)li.nCols, (unsigned)li.nRows,
                      (unsigned)li.nBands, (unsigned)li.numValid, li.blobSize, (unsigned)li.nMasks, (unsigned)li.nDepth,
                      (unsigned)li.nUsesNoData};
    for (int i = 0; i < infoArraySize; i++) infoArray[i] = i < 11 ? v[i] : 0;
  }
  if (dataRangeArray) {
    int noData = li.nDepth > 1 && li.nUsesNoData > 0;
    double v[3] = {noData ? -1 : li.zMin, noData ? -1 : li.zMax, li.maxZErr};
    for (int i = 0; i < dataRangeArraySize; i++) dataRangeArray[i] = i < 3 ? v[i] : 0;
  }
  return LO_OK;
}

unsigned lo_getDataRanges(const unsigned char* blob, unsigned blobSize, int nDepth, int nBands, double* mins, double* maxs) {
  if (!blob || !blobSize || !mins || !maxs || nDepth <= 0 || nBands <= 0) return LO_WRONG_PARAM;
  lo_info li;
  return blob_info(blob, blobSize, &li, mins, maxs, (size_t)nDepth * (size_t)nBands);
}

unsigned lo_decode_4D(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                      int nCols, int nRows, int nBands, unsigned dt, void* data, unsigned char* usesNoData, double* noDataValues);
unsigned lo_decode(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                   int nCols, int nRows, int nBands, unsigned dt, void* data) {
  return lo_decode_4D(blob, blobSize, nMasks, validBytes, nDepth, nCols, nRows, nBands, dt, data, NULL, NULL);
}
unsigned lo_decode_4D(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                      int nCols, int nRows, int nBands, unsigned dt, void* data, unsigned char* usesNoData, double* noDataValues) {
  if (!blob || !blobSize || !data || dt >= LO_UNDEFINED || nDepth <= 0 || nCols <= 0 || nRows <= 0 || nBands <= 0) return LO_WRONG_PARAM;
  if (!(nMasks == 0 || nMasks == 1 || nMasks == nBands) || (nMasks > 0 && !validBytes)) return LO_WRONG_PARAM;
  if (!dims_ok(nDepth, nCols, nRows, kTypeSize[dt])) return LO_DIMS_TOO_LARGE;
  lo_info li;
  unsigned e = blob_info(blob, blobSize, &li, NULL, NULL, 0);
  if (e) return e;
  if (nMasks < li.nMasks || nBands > li.nBands) return LO_WRONG_PARAM;    /* Lerc.cpp:423-428 */
  int wantNoData = li.nUsesNoData && nDepth > 1;
  if (wantNoData) {                                                        /* Lerc.cpp:431-445 */
    if (!usesNoData || !noDataValues) return LO_HAS_NODATA;
    memset(usesNoData, 0, (size_t)nBands); memset(noDataValues, 0, (size_t)nBands * sizeof(double));
  }
  if ((unsigned)li.dt != dt) return LO_FAILED;   /* deviation: the reference reinterprets; see DESIGN.md */
  return kDec[dt](blob, blobSize, nMasks, validBytes, nDepth, nCols, nRows, nBands, data, wantNoData ? usesNoData : NULL, noDataValues);
}

unsigned lo_decodeToDouble_4D(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                              int nCols, int nRows, int nBands, double* data, unsigned char* usesNoData, double* noDataValues);
unsigned lo_decodeToDouble(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                           int nCols, int nRows, int nBands, double* data) {
  return lo_decodeToDouble_4D(blob, blobSize, nMasks, validBytes, nDepth, nCols, nRows, nBands, data, NULL, NULL);
}
unsigned lo_decodeToDouble_4D(const unsigned char* blob, unsigned blobSize, int nMasks, unsigned char* validBytes, int nDepth,
                              int nCols, int nRows, int nBands, double* data, unsigned char* usesNoData, double* noDataValues) {       /* Lerc_c_api_impl.cpp:256-304 */
  if (!blob || !blobSize || !data || nDepth <= 0 || nCols <= 0 || nRows <= 0 || nBands <= 0) return LO_WRONG_PARAM;
  if (!(nMasks == 0 || nMasks == 1 || nMasks == nBands) || (nMasks > 0 && !validBytes)) return LO_WRONG_PARAM;
  lo_info li;
  unsigned e = blob_info(blob, blobSize, &li, NULL, NULL, 0);
  if (e) return e;
  if (li.nDepth != nDepth || li.nCols != nCols || li.nRows != nRows || li.nBands != nBands) return LO_FAILED;
  size_t n = (size_t)nDepth * (size_t)nCols * (size_t)nRows * (size_t)nBands;
  if (li.dt == LO_DOUBLE) return lo_decode_4D(blob, blobSize, nMasks, validBytes, nDepth, nCols, nRows, nBands, LO_DOUBLE, data, usesNoData, noDataValues);
  u8* tmp = (u8*)data + n * (8 - (size_t)kTypeSize[li.dt]);   /* decode into the tail, widen front to back */
  e = lo_decode_4D(blob, blobSize, nMasks, validBytes, nDepth, nCols, nRows, nBands, (unsigned)li.dt, tmp, usesNoData, noDataValues);
  if (e) return e;
  for (size_t k = 0; k < n; k++) data[k] = read_offset(tmp + k * (size_t)kTypeSize[li.dt], li.dt);
  return LO_OK;
}

/* ---- the _4D entry points (Lerc_c_api_impl.cpp:33-93, :178-252 with the noData arguments) ---- */
unsigned lo_computeCompressedSize_4D(const void* data, unsigned dt, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                                     const unsigned char* validBytes, double maxZErr, unsigned* numBytes,
                                     const unsigned char* usesNoData, const double* noDataValues) {
  if (!numBytes) return LO_WRONG_PARAM;
  *numBytes = 0;
  unsigned w = 0;
  return encode_common4(data, -1, dt, nDepth, nCols, nRows, nBands, nMasks, validBytes, maxZErr, NULL, 0, &w, numBytes, 1, usesNoData, noDataValues);
}

unsigned lo_encode_4D(const void* data, unsigned dt, int nDepth, int nCols, int nRows, int nBands, int nMasks,
                      const unsigned char* validBytes, double maxZErr, unsigned char* out, unsigned outSize, unsigned* nWritten,
                      const unsigned char* usesNoData, const double* noDataValues) {
  if (!nWritten) return LO_WRONG_PARAM;
  *nWritten = 0;
  unsigned need = 0;
  return encode_common4(data, -1, dt, nDepth, nCols, nRows, nBands, nMasks, validBytes, maxZErr, out, outSize, nWritten, &need, 0, usesNoData, noDataValues);
}
