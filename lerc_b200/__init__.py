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
