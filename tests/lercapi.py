"""ctypes view of the Lerc C API (include/Lerc_c_api.h) for ANY library exporting it.

Used by the tests, bench.py and __graft_entry__.smoke() to drive three libraries through the
very same calls: the product (lerc_b200/libLerc.so.4, CUDA), the oracle restatement
(oracle/_build/liblerc_oracle.so, plain C, exports the same 12 symbols with an `lo_` prefix)
and the unmodified reference (oracle/_ref/libLerc_ref.so).
Mirrors the call sequence of the reference's own wrapper, OtherLanguages/Python/lerc/_lerc.py:277-790.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DT_NP = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.float32, np.float64]
DT_CODE = {np.dtype(t): i for i, t in enumerate(DT_NP)}

ERR = {0: "Ok", 1: "Failed", 2: "WrongParam", 3: "BufferTooSmall", 4: "NaN", 5: "HasNoData", 6: "DimensionsTooLarge"}


class LercLib:
    def __init__(self, path, prefix="lerc_"):
        self.path = path
        self.lib = C.CDLL(path)
        self.prefix = prefix
        u, i, d, p = C.c_uint, C.c_int, C.c_double, C.c_void_p
        sig = {
            "computeCompressedSize": [p, u, i, i, i, i, i, p, d, p],
            "encode": [p, u, i, i, i, i, i, p, d, p, u, p],
            "computeCompressedSizeForVersion": [p, i, u, i, i, i, i, i, p, d, p],
            "encodeForVersion": [p, i, u, i, i, i, i, i, p, d, p, u, p],
            "getBlobInfo": [p, u, p, p, i, i],
            "getDataRanges": [p, u, i, i, p, p],
            "decode": [p, u, i, p, i, i, i, i, u, p],
            "decodeToDouble": [p, u, i, p, i, i, i, i, p],
            "computeCompressedSize_4D": [p, u, i, i, i, i, i, p, d, p, p, p],
            "encode_4D": [p, u, i, i, i, i, i, p, d, p, u, p, p, p],
            "decode_4D": [p, u, i, p, i, i, i, i, u, p, p, p],
            "decodeToDouble_4D": [p, u, i, p, i, i, i, i, p, p, p],
        }
        self.f = {}
        for name, args in sig.items():
            fn = getattr(self.lib, prefix + name, None)
            if fn is None:
                continue
            fn.argtypes = args
            fn.restype = C.c_uint
            self.f[name] = fn

    # ---- helpers -------------------------------------------------------------------------
    @staticmethod
    def _shape(arr, n_depth, n_bands):
        """arr layout: [nBands?][nRows][nCols][nDepth?]"""
        a = np.ascontiguousarray(arr)
        sh = list(a.shape)
        if n_depth > 1:
            assert sh[-1] == n_depth
            sh = sh[:-1]
        if n_bands > 1:
            assert sh[0] == n_bands
            sh = sh[1:]
        assert len(sh) == 2, sh
        return a, sh[0], sh[1]

    def compute_size(self, arr, max_z_err, n_depth=1, n_bands=1, mask=None, version=None):
        a, n_rows, n_cols = self._shape(arr, n_depth, n_bands)
        n_masks, mp = self._mask(mask, n_bands, n_rows, n_cols)
        n = C.c_uint(0)
        if version is None:
            st = self.f["computeCompressedSize"](a.ctypes.data, DT_CODE[a.dtype], n_depth, n_cols, n_rows, n_bands,
                                                 n_masks, mp, max_z_err, C.addressof(n))
        else:
            st = self.f["computeCompressedSizeForVersion"](a.ctypes.data, version, DT_CODE[a.dtype], n_depth, n_cols,
                                                           n_rows, n_bands, n_masks, mp, max_z_err, C.addressof(n))
        return st, n.value

    @staticmethod
    def _mask(mask, n_bands, n_rows, n_cols):
        if mask is None:
            return 0, None
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        n_masks = 1 if m.ndim == 2 else m.shape[0]
        assert m.size == n_masks * n_rows * n_cols
        LercLib._keep = m
        return n_masks, m.ctypes.data

    def encode(self, arr, max_z_err, n_depth=1, n_bands=1, mask=None, buf_size=None, version=None):
        """returns (status, blob bytes, whole zero-filled output buffer as np.uint8)"""
        a, n_rows, n_cols = self._shape(arr, n_depth, n_bands)
        n_masks, mp = self._mask(mask, n_bands, n_rows, n_cols)
        if buf_size is None:
            buf_size = int(a.nbytes * 1.1) + 4096 + (n_rows * n_cols * n_bands) // 4
        out = np.full(buf_size, 0xAB, dtype=np.uint8)  # the API must zero-fill it (Lerc.cpp:374)
        n = C.c_uint(0)
        if version is None:
            st = self.f["encode"](a.ctypes.data, DT_CODE[a.dtype], n_depth, n_cols, n_rows, n_bands, n_masks, mp,
                                  max_z_err, out.ctypes.data, buf_size, C.addressof(n))
        else:
            st = self.f["encodeForVersion"](a.ctypes.data, version, DT_CODE[a.dtype], n_depth, n_cols, n_rows, n_bands,
                                            n_masks, mp, max_z_err, out.ctypes.data, buf_size, C.addressof(n))
        return st, out[: n.value].tobytes(), out

    # ---- the _4D calls: per-band noData values (Lerc_c_api.h:295-380) -------------------------------
    def encode_4d(self, arr, max_z_err, n_depth=1, n_bands=1, mask=None, uses_no_data=None, no_data=None, buf_size=None, size_only=False):
        """returns (status, blob bytes) or, with size_only, (status, byte count)"""
        a, n_rows, n_cols = self._shape(arr, n_depth, n_bands)
        n_masks, mp = self._mask(mask, n_bands, n_rows, n_cols)
        u = None if uses_no_data is None else np.ascontiguousarray(uses_no_data, dtype=np.uint8)
        v = None if no_data is None else np.ascontiguousarray(no_data, dtype=np.float64)
        up, vp = (None if u is None else u.ctypes.data), (None if v is None else v.ctypes.data)
        n = C.c_uint(0)
        if size_only:
            st = self.f["computeCompressedSize_4D"](a.ctypes.data, DT_CODE[a.dtype], n_depth, n_cols, n_rows, n_bands, n_masks, mp, max_z_err,
                                                    C.addressof(n), up, vp)
            return st, n.value
        if buf_size is None:
            buf_size = int(a.nbytes * 1.1) + 4096 + (n_rows * n_cols * n_bands) // 4
        out = np.full(buf_size, 0xAB, dtype=np.uint8)
        st = self.f["encode_4D"](a.ctypes.data, DT_CODE[a.dtype], n_depth, n_cols, n_rows, n_bands, n_masks, mp, max_z_err,
                                 out.ctypes.data, buf_size, C.addressof(n), up, vp)
        return st, out[: n.value].tobytes()

    def decode_4d(self, blob, n_masks=None, to_double=False, want_no_data=True):
        """returns (status, data [nBands][nRows][nCols][nDepth], mask or None, usesNoData[nBands], noDataValues[nBands])"""
        b = np.frombuffer(blob, dtype=np.uint8)
        st, info = self.blob_info(blob)
        if st:
            return st, None, None, None, None
        nb, nr, nc, nd = info["nBands"], info["nRows"], info["nCols"], info["nDepth"]
        if n_masks is None:
            n_masks = info["nMasks"]
        dt = info["dataType"]
        data = np.full((nb, nr, nc, nd), 0x5A, dtype=np.float64 if to_double else DT_NP[dt])
        mask = np.full((n_masks, nr, nc), 7, dtype=np.uint8) if n_masks > 0 else None
        mp = mask.ctypes.data if mask is not None else None
        uses = np.full(nb, 9, dtype=np.uint8)
        vals = np.full(nb, -7.0, dtype=np.float64)
        up, vp = (uses.ctypes.data, vals.ctypes.data) if want_no_data else (None, None)
        if to_double:
            st = self.f["decodeToDouble_4D"](b.ctypes.data, b.size, n_masks, mp, nd, nc, nr, nb, data.ctypes.data, up, vp)
        else:
            st = self.f["decode_4D"](b.ctypes.data, b.size, n_masks, mp, nd, nc, nr, nb, dt, data.ctypes.data, up, vp)
        return st, data, mask, uses, vals

    def blob_info(self, blob):
        b = np.frombuffer(blob, dtype=np.uint8)
        info = np.zeros(11, dtype=np.uint32)
        rng = np.zeros(3, dtype=np.float64)
        st = self.f["getBlobInfo"](b.ctypes.data, b.size, info.ctypes.data, rng.ctypes.data, 11, 3)
        keys = ["version", "dataType", "nDepth", "nCols", "nRows", "nBands", "nValidPixels", "blobSize", "nMasks",
                "nDepth2", "nUsesNoDataValue"]
        d = {k: int(v) for k, v in zip(keys, info)}
        d.update(zMin=float(rng[0]), zMax=float(rng[1]), maxZErrUsed=float(rng[2]))
        return st, d

    def data_ranges(self, blob, n_depth, n_bands):
        b = np.frombuffer(blob, dtype=np.uint8)
        mins = np.zeros(n_depth * n_bands)
        maxs = np.zeros(n_depth * n_bands)
        st = self.f["getDataRanges"](b.ctypes.data, b.size, n_depth, n_bands, mins.ctypes.data, maxs.ctypes.data)
        return st, mins, maxs

    def decode(self, blob, n_masks=None, info=None, to_double=False):
        """returns (status, data array [nBands][nRows][nCols][nDepth], mask array or None)"""
        b = np.frombuffer(blob, dtype=np.uint8)
        if info is None:
            st, info = self.blob_info(blob)
            if st:
                return st, None, None
        nb, nr, nc, nd = info["nBands"], info["nRows"], info["nCols"], info["nDepth"]
        if n_masks is None:
            n_masks = info["nMasks"]
        dt = info["dataType"]
        data = np.full((nb, nr, nc, nd), 0x5A, dtype=np.float64 if to_double else DT_NP[dt])
        mask = np.full((n_masks, nr, nc), 7, dtype=np.uint8) if n_masks > 0 else None
        mp = mask.ctypes.data if mask is not None else None
        if to_double:
            st = self.f["decodeToDouble"](b.ctypes.data, b.size, n_masks, mp, nd, nc, nr, nb, data.ctypes.data)
        else:
            st = self.f["decode"](b.ctypes.data, b.size, n_masks, mp, nd, nc, nr, nb, dt, data.ctypes.data)
        return st, data, mask


def _first(*paths):
    for p in paths:
        if os.path.exists(p):
            return p
    return None


def ref_lib():
    p = _first(os.path.join(ROOT, "oracle", "_ref", "libLerc_ref.so"))
    return LercLib(p) if p else None


def oracle_lib(fpl_encoder=True):
    """fpl_encoder=False: float rasters at maxZError 0 are written without the lossless float codec (raw tiling / one sweep), as a
    build of the reference without fpl_*.cpp would -- a test switch of the oracle, not used by the suite"""
    p = _first(os.path.join(ROOT, "oracle", "_build", "liblerc_oracle.so"))
    if not p:
        return None
    lib = LercLib(p, prefix="lo_")
    lib.lib.lo_fpl_encoder(1 if fpl_encoder else 0)
    return lib


def product_lib():
    if os.environ.get("LERC_B200_SIM") == "1":       # development only (tools/cusim/README.md): never set by the driver, the gpu tests or bench.py
        return LercLib(os.path.join(ROOT, "tools", "cusim", os.environ.get("CUSIM_BUILD_DIR", "_build"), "libLerc_sim.so"))
    p = _first(os.path.join(ROOT, "lerc_b200", "libLerc.so.4"))
    return LercLib(p) if p else None


# ---- tile batch extension (include/lerc_b200.h) ---------------------------------------------------------
def tiles_api(lib):
    """ctypes signatures of lerc_b200_encodeTiles / decodeTiles / tilesMaxBytes on a loaded product library."""
    u, i, d, p, ull = C.c_uint, C.c_int, C.c_double, C.c_void_p, C.c_ulonglong
    lib.lib.lerc_b200_tilesMaxBytes.argtypes = [u, i, i, i, i]
    lib.lib.lerc_b200_tilesMaxBytes.restype = ull
    lib.lib.lerc_b200_encodeTiles.argtypes = [p, u, i, i, i, i, d, p, ull, p, p]
    lib.lib.lerc_b200_encodeTiles.restype = u
    lib.lib.lerc_b200_decodeTiles.argtypes = [p, ull, p, u, i, i, i, i, p]
    lib.lib.lerc_b200_decodeTiles.restype = u
    return lib.lib


def encode_tiles(lib, raster, tile_rows, tile_cols, max_z_err, buf_size=None):
    """host buffers through lerc_b200_encodeTiles: returns (status, list of blobs, offsets)"""
    api = tiles_api(lib)
    a = np.ascontiguousarray(raster)
    n_rows, n_cols = a.shape
    n_tiles = ((n_rows + tile_rows - 1) // tile_rows) * ((n_cols + tile_cols - 1) // tile_cols)
    if buf_size is None:
        buf_size = int(api.lerc_b200_tilesMaxBytes(DT_CODE[a.dtype], n_cols, n_rows, tile_cols, tile_rows))
    out = np.full(buf_size, 0xAB, dtype=np.uint8)
    off = np.zeros(n_tiles + 1, dtype=np.uint64)
    n = C.c_ulonglong(0)
    st = api.lerc_b200_encodeTiles(a.ctypes.data, DT_CODE[a.dtype], n_cols, n_rows, tile_cols, tile_rows, max_z_err,
                                   out.ctypes.data, buf_size, off.ctypes.data, C.addressof(n))
    if st:
        return st, None, None
    assert int(off[-1]) == n.value
    return st, [out[int(off[t]):int(off[t + 1])].tobytes() for t in range(n_tiles)], off


def decode_tiles(lib, blobs, dtype, n_rows, n_cols, tile_rows, tile_cols):
    api = tiles_api(lib)
    buf = np.frombuffer(b"".join(blobs) + b"\0" * 16, dtype=np.uint8)
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in blobs])
    out = np.full((n_rows, n_cols), 0x5A, dtype=dtype)
    st = api.lerc_b200_decodeTiles(buf.ctypes.data, int(off[-1]), off.ctypes.data, DT_CODE[np.dtype(dtype)], n_cols, n_rows,
                                   tile_cols, tile_rows, out.ctypes.data)
    return st, out


def tile_windows(n_rows, n_cols, tile_rows, tile_cols):
    for y in range(0, n_rows, tile_rows):
        for x in range(0, n_cols, tile_cols):
            yield slice(y, min(y + tile_rows, n_rows)), slice(x, min(x + tile_cols, n_cols))


def fletcher32(data):
    """Lerc2.cpp:1012-1064 over a bytes-like object (numpy; big-endian 16-bit words, sums start at 0xffff, odd tail byte << 8)"""
    b = np.frombuffer(bytes(data), dtype=np.uint8).astype(np.uint64)
    n = len(b) // 2
    w = (b[0:2 * n:2] << np.uint64(8)) + b[1:2 * n:2]
    if len(b) & 1:
        w = np.append(w, b[-1] << np.uint64(8))
    # sum1 = 0xffff + sum(w), sum2 = sum of the running sum1 values; both mod 65535 with 0 written as 0xffff (end-around carry)
    k = np.arange(len(w), 0, -1, dtype=np.uint64)
    s1 = (0xffff + int(w.sum())) % 65535
    s2 = (0xffff + 0xffff * len(w) + int((w * k % np.uint64(65535)).sum())) % 65535
    s1 = s1 or 0xffff
    s2 = s2 or 0xffff
    return (s2 << 16) | s1


def fpl_normalize(blob):
    """The reference leaves the read-ahead word behind every Huffman-coded byte plane of its lossless float codec uninitialised
    (fpl_EsriHuffman.cpp:403-448: malloc'ed, never written, copied into the blob).  Returns the blob with those 4 bytes per plane
    zeroed and the band checksums redone, so that reference-made blobs can be compared byte for byte; any other blob is returned
    unchanged (v6 float bands only)."""
    import struct
    b = bytearray(blob)
    pos = 0
    while pos + 90 <= len(b) and b[pos:pos + 6] == b"Lerc2 ":
        version = struct.unpack_from("<i", b, pos + 6)[0]
        if version != 6:
            break
        n_rows, n_cols, n_depth, n_valid, _mb, blob_size, dt, _more = struct.unpack_from("<8i", b, pos + 14)
        z_min, z_max = struct.unpack_from("<2d", b, pos + 58)
        p = pos + 90
        p += 4 + struct.unpack_from("<i", b, p)[0]
        if dt in (6, 7) and n_valid > 0 and z_min != z_max:
            sz = 4 if dt == 6 else 8
            ranges = bytes(b[p:p + 2 * sz * n_depth])
            p += 2 * sz * n_depth
            if ranges[:sz * n_depth] != ranges[sz * n_depth:] and b[p] == 0 and b[p + 1] == 3 and struct.unpack_from("<d", b, pos + 50)[0] == 0:
                p += 3
                for _ in range(sz):
                    size = struct.unpack_from("<I", b, p + 2)[0]
                    if b[p + 6] == 0:
                        b[p + 6 + size - 4:p + 6 + size] = b"\0\0\0\0"
                    p += 6 + size
                assert p == pos + blob_size
                struct.pack_into("<I", b, pos + 10, fletcher32(b[pos + 14:pos + blob_size]))
        pos += blob_size
    return bytes(b)
