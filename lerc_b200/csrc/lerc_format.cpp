// lerc_format.cpp -- host-only pieces of the Lerc2 byte stream: band header, multi-band header
// walk (lerc_getBlobInfo / lerc_getDataRanges) and the 256-symbol Huffman code table.
// These are O(header) / O(256) and stay on the CPU by design.
// Reference citations are relative to /root/reference/src/LercLib.
#include "lerc_internal.h"
#include <cstring>
#include <climits>
#include <cfloat>
#include <queue>
#include <algorithm>

namespace lerc {

// ------------------------------------------------------------------------------------------------
// header                                                                         Lerc2.cpp:710-917

int headerBytes(int v) {
  int n = 6 + 4;                       // "Lerc2 " + version
  if (v >= 3) n += 4;                  // checksum
  n += 4 * (v >= 4 ? 7 : 6);           // nRows nCols [nDepth] numValid mbSize blobSize dt
  if (v >= 6) n += 4 + 4;              // nBlobsMore + 4 flag bytes
  n += 8 * (v >= 6 ? 5 : 3);           // maxZErr zMin zMax [noData noDataOrig]
  return n;
}

namespace {
struct Writer {
  uint8_t* p;
  template <class V> void put(V v) { std::memcpy(p, &v, sizeof v); p += sizeof v; }
};
struct Reader {
  const uint8_t* p;
  template <class V> V get() { V v; std::memcpy(&v, p, sizeof v); p += sizeof v; return v; }
};
}  // namespace

void writeHeader(uint8_t* dst, const HeaderInfo& h) {
  Writer w{dst};
  std::memcpy(w.p, "Lerc2 ", 6); w.p += 6;
  w.put<int32_t>(h.version);
  if (h.version >= 3) w.put<uint32_t>(h.checksum);
  w.put<int32_t>(h.nRows); w.put<int32_t>(h.nCols);
  if (h.version >= 4) w.put<int32_t>(h.nDepth);
  w.put<int32_t>(h.numValidPixel); w.put<int32_t>(h.microBlockSize); w.put<int32_t>(h.blobSize); w.put<int32_t>(h.dt);
  if (h.version >= 6) {
    w.put<int32_t>(h.nBlobsMore);
    w.put<uint8_t>(h.bPassNoDataValues); w.put<uint8_t>(h.bIsInt); w.put<uint8_t>(h.bReserved3); w.put<uint8_t>(h.bReserved4);
  }
  w.put<double>(h.maxZError); w.put<double>(h.zMin); w.put<double>(h.zMax);
  if (h.version >= 6) { w.put<double>(h.noDataVal); w.put<double>(h.noDataValOrig); }
}

bool readHeader(const uint8_t* src, size_t avail, HeaderInfo& h) {
  h = HeaderInfo();
  h.version = 0;
  if (!src || avail < 10 || std::memcmp(src, "Lerc2 ", 6) != 0) return false;
  Reader r{src + 6};
  h.version = r.get<int32_t>();
  if (h.version < 0 || h.version > 6) return false;             // written by a newer codec
  if (avail < (size_t)headerBytes(h.version)) return false;
  if (h.version >= 3) h.checksum = r.get<uint32_t>();
  h.nRows = r.get<int32_t>(); h.nCols = r.get<int32_t>();
  h.nDepth = h.version >= 4 ? r.get<int32_t>() : 1;
  h.numValidPixel = r.get<int32_t>(); h.microBlockSize = r.get<int32_t>(); h.blobSize = r.get<int32_t>();
  int dt = r.get<int32_t>();
  if (h.version >= 6) {
    h.nBlobsMore = r.get<int32_t>();
    h.bPassNoDataValues = r.get<uint8_t>(); h.bIsInt = r.get<uint8_t>(); h.bReserved3 = r.get<uint8_t>(); h.bReserved4 = r.get<uint8_t>();
  }
  h.maxZError = r.get<double>(); h.zMin = r.get<double>(); h.zMax = r.get<double>();
  if (h.version >= 6) { h.noDataVal = r.get<double>(); h.noDataValOrig = r.get<double>(); }
  if (h.nRows <= 0 || h.nCols <= 0 || h.nDepth <= 0 || h.numValidPixel < 0 || h.microBlockSize <= 0 || h.blobSize <= 0 ||
      dt < DT_Char || dt > DT_Double)
    return false;
  h.dt = dt;
  const uint64_t nPix = (uint64_t)h.nRows * (uint64_t)h.nCols, lim = (uint64_t)INT_MAX, bpp = (uint64_t)typeSize(dt);
  if (nPix > lim || (uint64_t)h.numValidPixel > nPix) return false;
  if (h.microBlockSize > 32 || bpp * (uint64_t)h.nDepth > lim || bpp * (uint64_t)h.nDepth * nPix > lim) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------
// multi-band header walk                                           Lerc.cpp:92-182, :1014-1042

bool ByteSource::fetch(size_t off, size_t len, void* dst) const {
  if (off > size || len > size - off) return false;
  if (!onDevice) { std::memcpy(dst, base + off, len); return true; }
  // on the call's stream (lerc_b200_set_stream: "ordered on the given stream like any other work"), else the legacy default stream
  auto d2h = [&](void* to, const uint8_t* from, size_t n) {
    if (!stream) return cudaMemcpy(to, from, n, cudaMemcpyDeviceToHost) == cudaSuccess;
    return cudaMemcpyAsync(to, from, n, cudaMemcpyDeviceToHost, stream) == cudaSuccess && cudaStreamSynchronize(stream) == cudaSuccess;
  };
  if (len > sizeof cache) return d2h(dst, base + off, len);
  if (!(off >= cacheOff && off + len <= cacheOff + cacheLen)) {
    const size_t take = size - off < sizeof cache ? size - off : sizeof cache;
    if (!d2h(cache, base + off, take)) { cacheLen = 0; return false; }
    cacheOff = off; cacheLen = take;
  }
  std::memcpy(dst, cache + (off - cacheOff), len);
  return true;
}

namespace {

bool bandHeader(const ByteSource& s, size_t off, HeaderInfo& h, bool& hasMask) {   // Lerc2.cpp:495-510
  uint8_t buf[96];
  size_t avail = s.size > off ? s.size - off : 0;
  size_t take = avail < sizeof buf ? avail : sizeof buf;
  if (take < 14 || !s.fetch(off, take, buf)) return false;
  if (!readHeader(buf, take, h)) return false;
  size_t hb = (size_t)headerBytes(h.version);
  if (take < hb + 4) return false;
  int32_t nm; std::memcpy(&nm, buf + hb, 4);
  if (nm < 0) return false;
  hasMask = nm > 0;
  return true;
}

double readTyped(const uint8_t* p, int dt) {
  switch (dt) {
    case DT_Char:   { int8_t v;   std::memcpy(&v, p, 1); return v; }
    case DT_Byte:   { return *p; }
    case DT_Short:  { int16_t v;  std::memcpy(&v, p, 2); return v; }
    case DT_UShort: { uint16_t v; std::memcpy(&v, p, 2); return v; }
    case DT_Int:    { int32_t v;  std::memcpy(&v, p, 4); return v; }
    case DT_UInt:   { uint32_t v; std::memcpy(&v, p, 4); return v; }
    case DT_Float:  { float v;    std::memcpy(&v, p, 4); return v; }
    default:        { double v;   std::memcpy(&v, p, 8); return v; }
  }
}

ErrCode bandRanges(const ByteSource& s, size_t off, int iBand, const HeaderInfo& h, double* mins, double* maxs, size_t nElem) {
  const int nDepth = h.nDepth;
  if (nElem < ((size_t)iBand + 1) * (size_t)nDepth) return BufferTooSmall;
  if (nDepth == 1) { mins[iBand] = h.zMin; maxs[iBand] = h.zMax; return Ok; }
  if (h.bPassNoDataValues) return HasNoData;
  if (h.version < 4) return Failed;
  double* mn = mins + (size_t)iBand * nDepth;
  double* mx = maxs + (size_t)iBand * nDepth;
  if (h.numValidPixel == 0) { std::fill(mn, mn + nDepth, 0.0); std::fill(mx, mx + nDepth, 0.0); return Ok; }
  if (h.zMin == h.zMax) { std::fill(mn, mn + nDepth, h.zMin); std::fill(mx, mx + nDepth, h.zMin); return Ok; }
  size_t pos = off + (size_t)headerBytes(h.version);
  int32_t nm = 0;
  if (!s.fetch(pos, 4, &nm) || nm < 0) return Failed;
  pos += 4 + (size_t)nm;                                 // ranges follow the mask (Lerc2.cpp:415-426)
  const size_t ts = (size_t)typeSize(h.dt), len = ts * (size_t)nDepth;
  std::vector<uint8_t> buf(2 * len);
  if (!s.fetch(pos, 2 * len, buf.data())) return Failed;
  for (int i = 0; i < nDepth; i++) {
    mn[i] = readTyped(buf.data() + ts * (size_t)i, h.dt);
    mx[i] = readTyped(buf.data() + len + ts * (size_t)i, h.dt);
  }
  return Ok;
}

}  // namespace

ErrCode getBlobInfo(const ByteSource& s, BlobInfo& li, double* mins, double* maxs, size_t nElem) {
  li = BlobInfo();
  HeaderInfo h;
  bool hasMask = false;
  int nMasks = 0;
  if (!bandHeader(s, 0, h, hasMask)) return Failed;      // Lerc1 blobs are out of scope
  li.version = h.version; li.nDepth = h.nDepth; li.nCols = h.nCols; li.nRows = h.nRows;
  li.numValidPixel = h.numValidPixel; li.blobSize = (uint32_t)h.blobSize; li.dt = h.dt;
  li.zMin = h.zMin; li.zMax = h.zMax; li.maxZError = h.maxZError; li.nUsesNoDataValue = h.bPassNoDataValues ? 1 : 0;
  bool tryNext = h.version <= 5 || h.nBlobsMore > 0;
  if (hasMask || h.numValidPixel == 0) nMasks = 1;
  if (mins && maxs) { ErrCode e = bandRanges(s, 0, 0, h, mins, maxs, nElem); if (e != Ok) return e; }
  li.nBands = 1;
  if (li.blobSize > s.size) return Failed;               // truncated
  HeaderInfo g;
  while (tryNext && bandHeader(s, li.blobSize, g, hasMask)) {
    if (g.nDepth != li.nDepth || g.nCols != li.nCols || g.nRows != li.nRows || g.dt != li.dt) return Failed;
    tryNext = g.version <= 5 || g.nBlobsMore > 0;
    if (g.bPassNoDataValues) li.nUsesNoDataValue++;
    if (hasMask || g.numValidPixel != li.numValidPixel) nMasks = 2;
    const uint64_t sum = (uint64_t)li.blobSize + (uint64_t)g.blobSize;
    if (sum > (uint64_t)UINT_MAX || sum > (uint64_t)s.size) return Failed;
    li.zMin = std::min(li.zMin, g.zMin);
    li.zMax = std::max(li.zMax, g.zMax);
    li.maxZError = std::max(li.maxZError, g.maxZError);
    if (mins && maxs) { ErrCode e = bandRanges(s, li.blobSize, li.nBands, g, mins, maxs, nElem); if (e != Ok) return e; }
    li.blobSize += (uint32_t)g.blobSize;
    li.nBands++;
  }
  li.nMasks = nMasks > 1 ? li.nBands : nMasks;
  if (li.nUsesNoDataValue > 0) li.nUsesNoDataValue = li.nBands;
  return Ok;
}

// ------------------------------------------------------------------------------------------------
// Huffman code table (256 symbols)                                               Huffman.cpp:35-572

namespace {

int bitLength(uint32_t v) { int n = 0; while (n < 32 && (v >> n)) n++; return n; }
int wrapIdx(int i, int size) { return i < size ? i : i - size; }

struct PqNode {
  int weight;      // -count, so the rarest symbol is the priority queue's top (Huffman.h:90)
  int id;          // index into the node pool
  bool operator<(const PqNode& o) const { return weight < o.weight; }
};
struct TreeNode { int leaf, kid0, kid1; };

bool assignDepths(const std::vector<TreeNode>& pool, int root, int depth, uint16_t* len) {
  const TreeNode& n = pool[root];
  if (n.leaf >= 0) { len[n.leaf] = (uint16_t)depth; return true; }
  if (depth == 32) return false;                       // Huffman.h:91: codes longer than 32 bits are refused
  return assignDepths(pool, n.kid0, depth + 1, len) && assignDepths(pool, n.kid1, depth + 1, len);
}

// fixed-width LSB-first packing of the code-length array (BitStuffer2::EncodeSimple, BitStuffer2.cpp:35-75)
// Lerc2 v2 bit stuffing (BitStuffer2.cpp:292-425): values MSB-first inside little-endian uint32 words, the unused low bytes of
// the last word dropped by shifting that word down.  Stored byte that holds stream bit s (bit 7 - s % 8 of that byte), or -1.
static long v2ByteOfStreamBit(uint64_t s, uint32_t n, int nb) {
  const uint64_t total = (uint64_t)n * nb, nWords = (total + 31) / 32, w = s >> 5;
  const int bitsTail = (int)(total & 31), bytesTail = (bitsTail + 7) >> 3, drop = bytesTail > 0 ? 4 - bytesTail : 0;
  const int jj = 3 - (int)((s & 31) >> 3);                       // byte of the little-endian word
  const long k = (long)(4 * w) + jj - (w == nWords - 1 ? drop : 0);
  return k >= (long)(4 * w) ? k : -1;
}

size_t stuffSimple(uint8_t* dst, const uint32_t* v, uint32_t n, int version) {
  uint32_t mx = 0;
  for (uint32_t i = 0; i < n; i++) mx = std::max(mx, v[i]);
  const int nb = bitLength(mx), cb = n < 256 ? 1 : (n < 65536 ? 2 : 4);
  uint8_t* p = dst;
  *p++ = (uint8_t)(nb | ((cb == 4 ? 0 : 3 - cb) << 6));
  if (cb == 1) *p = (uint8_t)n; else if (cb == 2) { uint16_t s = (uint16_t)n; std::memcpy(p, &s, 2); } else std::memcpy(p, &n, 4);
  p += cb;
  if (nb > 0) {
    const size_t len = ((size_t)n * nb + 7) >> 3;
    std::memset(p, 0, len);
    if (version >= 3) {
      uint64_t bit = 0;
      for (uint32_t i = 0; i < n; i++, bit += nb) {
        uint64_t x = (uint64_t)v[i] << (bit & 7);
        for (size_t k = bit >> 3; x; k++, x >>= 8) p[k] |= (uint8_t)x;
      }
    } else {
      for (uint32_t i = 0; i < n; i++)
        for (int t = 0; t < nb; t++)
          if ((v[i] >> (nb - 1 - t)) & 1) { const long k = v2ByteOfStreamBit((uint64_t)i * nb + t, n, nb); if (k >= 0 && (size_t)k < len) p[k] |= (uint8_t)(0x80u >> (((uint64_t)i * nb + t) & 7)); }
    }
    p += len;
  }
  return (size_t)(p - dst);
}

size_t unstuffSimple(const uint8_t* src, size_t avail, uint32_t* v, uint32_t expect, int version) {   // BitStuffer2.cpp:159-196
  if (avail < 1) return 0;
  const uint8_t b = src[0];
  const int code = b >> 6, cb = code == 0 ? 4 : 3 - code, nb = b & 31;
  if ((b >> 5) & 1) return 0;                          // LUT mode is never used for the code lengths
  if (cb <= 0 || avail < 1 + (size_t)cb) return 0;
  uint32_t n = 0;
  if (cb == 1) n = src[1]; else if (cb == 2) { uint16_t s; std::memcpy(&s, src + 1, 2); n = s; } else std::memcpy(&n, src + 1, 4);
  if (n != expect) return 0;
  const size_t len = ((size_t)n * nb + 7) >> 3;
  if (avail < 1 + (size_t)cb + len) return 0;
  const uint8_t* p = src + 1 + cb;
  uint64_t bit = 0;
  const uint32_t mask = nb ? ((nb == 32) ? 0xffffffffu : ((1u << nb) - 1)) : 0;
  if (version >= 3) {
    for (uint32_t i = 0; i < n; i++, bit += nb) {
      uint64_t x = 0;
      const size_t k0 = bit >> 3;
      for (int k = 0; k < 5 && k0 + k < len; k++) x |= (uint64_t)p[k0 + k] << (8 * k);
      v[i] = (uint32_t)(x >> (bit & 7)) & mask;
    }
  } else {
    if (nb > 0 && n == 0) return 0;                    // BitUnStuff_Before_Lerc2v3 rejects an empty array (BitStuffer2.cpp:355)
    for (uint32_t i = 0; i < n; i++) {
      uint32_t x = 0;
      for (int t = 0; t < nb; t++) {
        const uint64_t sb = (uint64_t)i * nb + t;
        const long k = v2ByteOfStreamBit(sb, n, nb);
        const uint32_t b1 = (k >= 0 && (size_t)k < len) ? ((p[k] >> (7 - (sb & 7))) & 1u) : 0u;
        x = (x << 1) | b1;
      }
      v[i] = x;
    }
  }
  return 1 + (size_t)cb + len;
}

inline void putBitsMsb(uint8_t* base, uint64_t& bitPos, uint32_t val, int nBits) {      // Huffman.h:218-255
  for (int k = nBits - 1; k >= 0; k--, bitPos++)
    if ((val >> k) & 1) {
      const uint64_t word = bitPos >> 5; const int bit = 31 - (int)(bitPos & 31);
      base[word * 4 + (bit >> 3)] |= (uint8_t)(1u << (bit & 7));
    }
}
inline int getBitMsb(const uint8_t* base, uint64_t bitPos) {
  const uint64_t word = bitPos >> 5; const int bit = 31 - (int)(bitPos & 31);
  return (base[word * 4 + (bit >> 3)] >> (bit & 7)) & 1;
}

}  // namespace

bool HuffmanTable::buildFromHistogram(const int* histo) {
  std::fill(len, len + 256, (uint16_t)0);
  std::fill(code, code + 256, 0u);
  // The tie-breaking of equal counts is whatever std::priority_queue does (Huffman.cpp:40-61); using the
  // same container adaptor on the same libstdc++ reproduces the reference's code lengths.
  std::priority_queue<PqNode> pq;
  std::vector<TreeNode> pool;
  pool.reserve(512);
  for (int i = 0; i < 256; i++)
    if (histo[i] > 0) { pool.push_back({i, -1, -1}); pq.push({-histo[i], (int)pool.size() - 1}); }
  if (pq.size() < 2) return false;
  while (pq.size() > 1) {
    PqNode a = pq.top(); pq.pop();
    PqNode b = pq.top(); pq.pop();
    pool.push_back({-1, a.id, b.id});
    pq.push({a.weight + b.weight, (int)pool.size() - 1});
  }
  if (!assignDepths(pool, pq.top().id, 0, len)) return false;
  // canonical codes: sort by (length desc, symbol asc); walking that order the code counts up and is
  // shifted right whenever the length drops (Huffman.cpp:541-572)
  std::vector<std::pair<int, int>> order;
  for (int i = 0; i < 256; i++) if (len[i]) order.push_back({len[i] * 256 - i, i});
  std::sort(order.begin(), order.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first > y.first; });
  int curLen = len[order[0].second];
  uint32_t c = 0;
  for (auto& o : order) {
    const int s = o.second, d = curLen - len[s];
    c >>= d; curLen -= d;
    code[s] = c++;
  }
  return true;
}

bool HuffmanTable::range(int& i0, int& i1, int& maxLen) const {
  const int size = 256;
  int a = 0, b = size - 1;
  while (a < size && len[a] == 0) a++;
  while (b >= 0 && len[b] == 0) b--;
  if (b + 1 <= a) return false;
  i0 = a; i1 = b + 1;
  int bestStart = 0, bestLen = 0, j = 0;        // longest stretch of unused symbols; wrap around it if that is shorter
  while (j < size) {
    while (j < size && len[j] > 0) j++;
    const int k0 = j;
    while (j < size && len[j] == 0) j++;
    if (j - k0 > bestLen) { bestStart = k0; bestLen = j - k0; }
  }
  if (size - bestLen < i1 - i0) { i0 = bestStart + bestLen; i1 = bestStart + size; }
  if (i1 <= i0) return false;
  int mx = 0;
  for (int i = i0; i < i1; i++) mx = std::max<int>(mx, len[wrapIdx(i, size)]);
  if (mx <= 0 || mx > 32) return false;
  maxLen = mx;
  return true;
}

bool HuffmanTable::tableBytes(int& nBytes) const {
  int i0, i1, maxLen;
  if (!range(i0, i1, maxLen)) return false;
  int sum = 0;
  for (int i = i0; i < i1; i++) sum += len[wrapIdx(i, 256)];
  const uint32_t n = (uint32_t)(i1 - i0);
  const int lensBytes = 1 + (n < 256 ? 1 : 2) + (int)(((size_t)n * bitLength((uint32_t)maxLen) + 7) >> 3);
  nBytes = 16 + lensBytes + 4 * ((((sum + 7) >> 3) + 3) >> 2);
  return true;
}

bool HuffmanTable::totalBytes(const int* histo, int& nBytes) const {
  int tb;
  if (!tableBytes(tb)) return false;
  int64_t bits = 0, elems = 0;
  for (int i = 0; i < 256; i++) if (histo[i] > 0) { bits += (int64_t)histo[i] * len[i]; elems += histo[i]; }
  // The reference accumulates the bit count in a 32-bit int (Huffman.cpp:94-107); past 2^31 bits it
  // overflows and falls back to tiling (Lerc2.cpp:294-295).  We make that explicit.
  if (elems == 0 || bits >= ((int64_t)1 << 31)) return false;
  const int64_t total = tb + 4 * (((((bits + 7) >> 3) + 3) >> 2) + 1);
  if (total > INT_MAX) return false;
  nBytes = (int)total;
  return true;
}

size_t HuffmanTable::write(uint8_t* dst, int version) const {
  int i0, i1, maxLen;
  if (!range(i0, i1, maxLen)) return 0;
  Writer w{dst};
  w.put<int32_t>(4); w.put<int32_t>(256); w.put<int32_t>(i0); w.put<int32_t>(i1);
  uint32_t lens[512];
  for (int i = i0; i < i1; i++) lens[i - i0] = len[wrapIdx(i, 256)];
  w.p += stuffSimple(w.p, lens, (uint32_t)(i1 - i0), version);
  uint64_t bitPos = 0;
  for (int i = i0; i < i1; i++) { const int k = wrapIdx(i, 256); if (len[k]) putBitsMsb(w.p, bitPos, code[k], len[k]); }
  w.p += 4 * ((bitPos + 31) >> 5);
  return (size_t)(w.p - dst);
}

size_t HuffmanTable::read(const uint8_t* src, size_t avail, int version) {
  if (avail < 16) return 0;
  Reader r{src};
  const int ver = r.get<int32_t>(), size = r.get<int32_t>(), i0 = r.get<int32_t>(), i1 = r.get<int32_t>();
  if (ver < 2 || i0 >= i1 || i0 < 0 || size < 0 || size > 256 || i1 - i0 > 512) return 0;
  if (wrapIdx(i0, size) >= size || wrapIdx(i1 - 1, size) >= size) return 0;
  uint32_t lens[512];
  const size_t used = unstuffSimple(r.p, avail - 16, lens, (uint32_t)(i1 - i0), version);
  if (!used) return 0;
  r.p += used;
  std::fill(len, len + 256, (uint16_t)0);
  std::fill(code, code + 256, 0u);
  uint64_t bits = 0;
  for (int i = i0; i < i1; i++) {
    if (lens[i - i0] > 32) return 0;
    len[wrapIdx(i, size)] = (uint16_t)lens[i - i0];
    bits += lens[i - i0];
  }
  const size_t bytes = 4 * (size_t)((bits + 31) >> 5);
  if (avail - 16 - used < bytes) return 0;
  uint64_t pos = 0;
  for (int i = i0; i < i1; i++) {
    const int k = wrapIdx(i, size);
    uint32_t c = 0;
    for (int b = 0; b < len[k]; b++) c = (c << 1) | (uint32_t)getBitMsb(r.p, pos++);
    code[k] = c;
  }
  return 16 + used + bytes;
}

}  // namespace lerc
