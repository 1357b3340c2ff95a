"""lerc_b200 -- Blackwell-native LERC (Lerc2) encode/decode behind the unchanged Esri/lerc C API.

The product is the shared library lerc_b200/libLerc.so.4 (hand-written CUDA kernels for sm_100a plus a
C++ host layer; sources in lerc_b200/csrc).  This package only locates and loads it and offers the same
Python surface as the reference's ctypes wrapper (OtherLanguages/Python/lerc/_lerc.py): see lerc_b200.api.
There is no CPU fallback: importing works without a GPU (so the build can be checked), every codec call
fails with status 1 (Failed) when no CUDA device is present.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libLerc.so.4")

REQUIRED_SYMBOLS = [
    "lerc_computeCompressedSize", "lerc_encode", "lerc_computeCompressedSizeForVersion", "lerc_encodeForVersion",
    "lerc_getBlobInfo", "lerc_getDataRanges", "lerc_decode", "lerc_decodeToDouble",
    "lerc_computeCompressedSize_4D", "lerc_encode_4D", "lerc_decode_4D", "lerc_decodeToDouble_4D",
    "lerc_b200_set_stream", "lerc_b200_get_stats", "lerc_b200_version", "lerc_b200_profile", "lerc_b200_get_profile",
    "lerc_b200_tilesMaxBytes", "lerc_b200_encodeTiles", "lerc_b200_decodeTiles",
]


def load_library():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python __graft_entry__.py` (or make -C lerc_b200/csrc)")
    lib = ctypes.CDLL(LIB_PATH)
    missing = [s for s in REQUIRED_SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise ImportError(f"{LIB_PATH} lacks symbols: {missing}")
    return lib


_lib = load_library()
_lib.lerc_b200_version.restype = ctypes.c_char_p
__version__ = _lib.lerc_b200_version().decode()


def stats():
    """(kernel launches, encode calls, decode calls, fast-path encodes, fast-path decodes) since load."""
    out = (ctypes.c_ulonglong * 5)()
    _lib.lerc_b200_get_stats(out, 5)
    return tuple(int(v) for v in out)


def set_stream(cuda_stream_ptr, enable=True):
    """Route this thread's lerc_* calls onto the given cudaStream_t (integer handle, e.g.
    torch.cuda.current_stream().cuda_stream)."""
    _lib.lerc_b200_set_stream(ctypes.c_void_p(cuda_stream_ptr), 1 if enable else 0)


def profile(enable=True):
    """Bracket every kernel launch with CUDA events (see kernel_times())."""
    _lib.lerc_b200_profile(1 if enable else 0)


def kernel_times(reset=True):
    """{kernel name: (launches, total milliseconds)} accumulated while profile(True) was on."""
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.lerc_b200_get_profile(buf, len(buf), 1 if reset else 0)
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split("\t")
        out[name] = (int(cnt), float(ms))
    return out


# ---- tile batch (include/lerc_b200.h): thin wrappers over lerc_b200_encodeTiles / lerc_b200_decodeTiles ------------------------------
_DT_CODES = {"int8": 0, "uint8": 1, "int16": 2, "uint16": 3, "int32": 4, "uint32": 5, "float32": 6, "float64": 7}
_lib.lerc_b200_tilesMaxBytes.restype = ctypes.c_ulonglong
_lib.lerc_b200_tilesMaxBytes.argtypes = [ctypes.c_uint] + [ctypes.c_int] * 4
_lib.lerc_b200_encodeTiles.restype = ctypes.c_uint
_lib.lerc_b200_encodeTiles.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                       ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p]
_lib.lerc_b200_decodeTiles.restype = ctypes.c_uint
_lib.lerc_b200_decodeTiles.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_uint, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p]


def _ptr(buf):
    """address of a numpy array (host) or a torch tensor (host or CUDA): the library accepts both kinds of pointers"""
    if hasattr(buf, "data_ptr"):
        return buf.data_ptr()
    return buf.ctypes.data


def _dtype_code(buf):
    name = str(buf.dtype).replace("torch.", "")
    return _DT_CODES[name]


def tiles_max_bytes(dtype_code, n_rows, n_cols, tile_rows, tile_cols):
    """an output size that always suffices for encode_tiles (lerc_b200_tilesMaxBytes)"""
    return int(_lib.lerc_b200_tilesMaxBytes(dtype_code, n_cols, n_rows, tile_cols, tile_rows))


def encode_tiles(raster, tile_rows, tile_cols, max_z_err, out, offsets):
    """raster: contiguous [n_rows][n_cols] numpy array or torch tensor (host or CUDA); out: uint8 buffer of at least
    tiles_max_bytes(...) bytes; offsets: 64-bit integer buffer with n_tiles + 1 entries.  Every tile_rows x tile_cols window becomes
    its own standard Lerc2 blob, blob t = out[offsets[t]:offsets[t + 1]].  Returns (lerc_status, bytes written)."""
    n_rows, n_cols = int(raster.shape[0]), int(raster.shape[1])
    n = ctypes.c_ulonglong(0)
    cap = int(out.numel() if hasattr(out, "numel") else out.size)
    st = _lib.lerc_b200_encodeTiles(_ptr(raster), _dtype_code(raster), n_cols, n_rows, tile_cols, tile_rows, float(max_z_err),
                                    _ptr(out), cap, _ptr(offsets), ctypes.addressof(n))
    return int(st), int(n.value)


def decode_tiles(blobs, n_bytes, offsets, out, tile_rows, tile_cols):
    """inverse of encode_tiles: decodes the n_tiles blobs addressed by `offsets` into the [n_rows][n_cols] buffer `out`.
    Returns the lerc_status."""
    n_rows, n_cols = int(out.shape[0]), int(out.shape[1])
    return int(_lib.lerc_b200_decodeTiles(_ptr(blobs), int(n_bytes), _ptr(offsets), _dtype_code(out), n_cols, n_rows, tile_cols, tile_rows, _ptr(out)))
