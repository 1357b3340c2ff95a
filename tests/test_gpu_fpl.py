"""GPU tests (-m gpu): the lossless float codec.  ENCODE (lerc_b200/csrc/lerc_fpl_encode.cuh): float rasters at maxZError 0 must
come out as the reference's blobs (fpl_ref.npz, up to the 4 uninitialised bytes the reference leaves per Huffman plane, see
lercapi.fpl_normalize) and as the oracle's.  DECODE: blobs written by the reference's lossless float codec (FPL, image mode IEM_DeltaDeltaHuffman,
fpl_*.cpp; product: lerc_b200/csrc/lerc_fpl_decode.cuh).  The blobs are the reference's own output for tests/cases.py:fpl_cases
(committed in tests/golden/fpl_ref.npz with the hash of what the reference decodes); the oracle's decoder is pinned to those
hashes by tests/test_oracle_vs_reference.py::test_fpl_blobs_decode_like_the_reference."""
import hashlib
import os

import numpy as np
import pytest

from cases import c2_raster, fpl_cases, fpl_encode_cases, fpl_fuzz_cases
from lercapi import ROOT, fpl_normalize, oracle_lib, product_lib, ref_lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = fpl_cases()
ENC_CASES = fpl_encode_cases()


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib(fpl_encoder=True)
    assert prod is not None and orc is not None
    return prod, orc


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fpl_blob_decodes_like_reference(libs, case):
    prod, orc = libs
    name, arr, kw = case
    g = np.load(os.path.join(GOLD, "fpl_ref.npz"))
    blob = g["blob_" + name].tobytes()
    st, data, mask = prod.decode(blob)
    assert st == 0
    h = hashlib.sha256(data.tobytes())
    if mask is not None:
        h.update(mask.tobytes())
    assert h.hexdigest() == str(g["hash_" + name]), "decoded pixels differ from the reference's"
    t_o, d_o, m_o = orc.decode(blob)
    assert t_o == 0 and np.array_equal(d_o.view(np.uint8), data.view(np.uint8))
    # lossless: the valid pixels are the input, bit for bit (NaN pixels were moved to the mask by the encoder)
    a = np.ascontiguousarray(arr).reshape(data.shape)
    for b in range(data.shape[0]):
        valid = np.ones(data.shape[1:3], bool) if mask is None else mask[b if mask.shape[0] > 1 else 0].astype(bool)
        assert np.array_equal(data[b][valid].view(np.uint8), a[b][valid].view(np.uint8))


def test_corrupted_fpl_blobs_fail_like_the_oracle(libs):
    """(the reference's FPL reader has no bound checks on the plane streams -- it accepts all of these and may read past the
    plane; the oracle and the product apply the 8-bit Huffman path's rules and agree with each other)"""
    prod, orc = libs
    g = np.load(os.path.join(GOLD, "fpl_ref.npz"))
    blob = bytearray(g["blob_f32_noisy"].tobytes())
    rng = np.random.default_rng(5)
    import ctypes as C
    fl = orc.lib.lo_fletcher32
    fl.restype = C.c_uint32
    fl.argtypes = [C.c_void_p, C.c_int]
    for _ in range(40):
        bad = bytearray(blob)
        k = int(rng.integers(110, len(bad) - 8))
        bad[k] ^= int(rng.integers(1, 256))
        buf = np.frombuffer(bytes(bad), np.uint8).copy()
        cs = fl(buf[14:].ctypes.data, len(bad) - 14)                       # repair the checksum so that the FPL parser is reached
        buf[10:14] = np.frombuffer(np.uint32(cs).tobytes(), np.uint8)
        s_o, d_o, _ = orc.decode(buf.tobytes())
        s_p, d_p, _ = prod.decode(buf.tobytes())
        assert (s_p == 0) == (s_o == 0), f"byte {k}: status {s_p} vs oracle {s_o}"
        if s_o == 0:
            assert np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fpl_encoder_writes_the_references_blob(libs, case):
    prod, orc = libs
    name, arr, kw = case
    g = np.load(os.path.join(GOLD, "fpl_ref.npz"))
    want = fpl_normalize(g["blob_" + name].tobytes())
    st, blob, whole = prod.encode(arr, 0.0, **kw)
    assert st == 0
    assert blob == want, f"len {len(blob)} vs the reference's {len(want)}"
    assert not whole[len(blob):].any()                                       # Lerc.cpp:374: the rest of the buffer stays zero
    assert prod.compute_size(arr, 0.0, **kw) == (0, len(want))
    assert prod.encode(arr, 0.0, buf_size=len(want) - 1, **kw)[0] == 3      # BufferTooSmall


@pytest.mark.parametrize("case", ENC_CASES, ids=[c[0] for c in ENC_CASES])
def test_fpl_encoder_matches_oracle(libs, case):
    prod, orc = libs
    name, arr, kw = case
    s_o, b_o, _ = orc.encode(arr, 0.0, **kw)
    s_p, b_p, _ = prod.encode(arr, 0.0, **kw)
    assert s_o == 0 and s_p == 0
    assert b_p == b_o, f"len {len(b_p)} vs {len(b_o)}"
    st, data, mask = prod.decode(b_p)
    t_o, d_o, m_o = orc.decode(b_p)
    assert st == 0 and t_o == 0 and np.array_equal(d_o.view(np.uint8), data.view(np.uint8))
    a = np.ascontiguousarray(arr).reshape(data.shape)
    valid = np.ones(data.shape[1:3], bool) if mask is None else mask[0].astype(bool)
    assert np.array_equal(data[0][valid].view(np.uint8), a[0][valid].view(np.uint8))


def test_fpl_encoder_fuzz(libs):
    """40 seeded random rasters (degenerate shapes, nDepth, masks, NaNs): status and blob equal the oracle's"""
    prod, orc = libs
    for name, arr, kw in fpl_fuzz_cases():
        s_o, b_o, _ = orc.encode(arr, 0.0, **kw)
        s_p, b_p, _ = prod.encode(arr, 0.0, **kw)
        assert s_p == s_o, name
        assert b_p == b_o, f"{name}: len {len(b_p or b'')} vs {len(b_o or b'')}"


def test_fpl_full_size_lossless_round_trip(libs):
    """BASELINE configs[1]'s raster (4096 x 4096 float32) at maxZError 0: the blob is the reference's (the oracle's where the
    reference library is not on the box), and decode(encode(x)) == x bit for bit"""
    prod, orc = libs
    img = c2_raster(4096, 4096)
    chk = ref_lib() or orc
    s_r, b_r, _ = chk.encode(img, 0.0)
    s_p, b_p, _ = prod.encode(img, 0.0)
    assert s_r == 0 and s_p == 0
    assert b_p == fpl_normalize(b_r), f"len {len(b_p)} vs {len(b_r)}"
    assert len(b_p) < 0.9 * img.nbytes                                       # the codec was taken (raw tiling would be > 64 MiB)
    st, data, _ = prod.decode(b_p)
    assert st == 0 and np.array_equal(data.reshape(img.shape).view(np.uint32), img.view(np.uint32))
