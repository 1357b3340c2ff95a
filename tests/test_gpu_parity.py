"""GPU tests (-m gpu): the CUDA product (lerc_b200/libLerc.so.4, called through the C ABI) against the oracle.

Bar (task section 3): bit-exact blobs and bit-exact decoded pixels for every pixel type, lossy float
included (same fp64 formula, no FMA contraction).  Sizes here are ones the scalar oracle finishes in
seconds; full BASELINE sizes are covered by size-independent properties in test_gpu_large.py.
"""
import os

import numpy as np
import pytest

from cases import all_cases, c2_raster, c4_raster
from lercapi import ROOT, oracle_lib, product_lib, ref_lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = all_cases()


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib()
    assert prod is not None, "lerc_b200/libLerc.so.4 missing"
    assert orc is not None, "oracle/_build/liblerc_oracle.so missing"
    return prod, orc


def _first_diff(a, b):
    n = min(len(a), len(b))
    x, y = np.frombuffer(a[:n], np.uint8), np.frombuffer(b[:n], np.uint8)
    d = np.nonzero(x != y)[0]
    return f"len {len(a)} vs {len(b)}, first diffs at {d[:6].tolist()}"


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_encode_matches_oracle(libs, case):
    prod, orc = libs
    name, arr, mz, kw = case
    s_o, b_o, _ = orc.encode(arr, mz, **kw)
    s_p, b_p, buf = prod.encode(arr, mz, **kw)
    assert s_p == s_o, f"status {s_p} vs oracle {s_o}"
    if s_o != 0:
        return
    assert b_p == b_o, _first_diff(b_p, b_o)
    assert not buf[len(b_p):].any(), "output buffer not zero-filled past the blob (Lerc.cpp:374)"
    st, n = prod.compute_size(arr, mz, **kw)
    assert st == 0 and n == len(b_o)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_decode_matches_oracle(libs, case):
    prod, orc = libs
    name, arr, mz, kw = case
    s_o, blob, _ = orc.encode(arr, mz, **kw)
    if s_o != 0:
        pytest.skip("oracle refuses this input")
    t_o, d_o, m_o = orc.decode(blob)
    t_p, d_p, m_p = prod.decode(blob)
    assert t_p == 0 and t_o == 0
    assert np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8)), "decoded pixels differ bitwise"
    assert np.array_equal(m_p, m_o)
    if mz == 0 and "nan" not in name:
        want = np.ascontiguousarray(arr).reshape(d_p.shape)
        valid = np.ones(d_p.shape, bool) if m_p is None else np.broadcast_to(m_p.reshape(m_p.shape[0], *d_p.shape[1:3], 1).astype(bool), d_p.shape) if m_p.shape[0] == d_p.shape[0] else np.broadcast_to(m_p.reshape(1, *d_p.shape[1:3], 1).astype(bool), d_p.shape)
        assert np.array_equal(d_p[valid], want[valid]), "lossless round trip broken"


@pytest.mark.parametrize("name", ["california_400_400_1_float", "bluemarble_256_256_3_byte", "js_sanity_30_20_3_byte"])
def test_golden_fixtures_decode(libs, name):
    prod, _ = libs
    blob = open(os.path.join(GOLD, name + ".lerc2"), "rb").read()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    st, data, mask = prod.decode(blob)
    assert st == 0
    assert np.array_equal(data.view(np.uint8), g["data"].view(np.uint8))
    if g["mask"].size:
        assert np.array_equal(mask, g["mask"])
    st, dbl, _ = prod.decode(blob, to_double=True)
    assert st == 0 and np.array_equal(dbl, g["data"].astype(np.float64))


def test_bluemarble_reencode_reproduces_huffman_stream(libs):
    """re-encoding the decoded bluemarble reproduces the reference's Huffman table and bit
    stream; our v6 blob must equal the oracle's (== reference's) v6 blob of the same pixels."""
    prod, orc = libs
    g = np.load(os.path.join(GOLD, "bluemarble_256_256_3_byte.npz"))
    data, mask = g["data"][..., 0], g["mask"][0]
    s_p, b_p, _ = prod.encode(data, 0, n_bands=3, mask=mask)
    s_o, b_o, _ = orc.encode(data, 0, n_bands=3, mask=mask)
    assert s_p == 0 and s_o == 0 and b_p == b_o, _first_diff(b_p, b_o)
    st, info = prod.blob_info(b_p)
    assert info["nBands"] == 3 and info["nMasks"] == 1


def test_error_codes(libs):
    prod, orc = libs
    a = np.ones((16, 16), np.float32)
    st, blob, _ = prod.encode(a * 2 + np.arange(16, dtype=np.float32), 0.01, buf_size=60)
    assert st == 3                                                # BufferTooSmall
    nan = np.full((4, 4, 2), 1.0, np.float32); nan[0, 0, 0] = np.nan
    assert prod.encode(nan, 0.01, n_depth=2)[0] == 4              # NaN
    st, blob, _ = prod.encode(a + np.arange(16, dtype=np.float32), 0.01)
    assert st == 0
    bad = bytearray(blob); bad[-1] ^= 0xFF
    assert prod.decode(bytes(bad))[0] == 1                        # checksum mismatch
    bad = bytearray(blob); bad[100] ^= 0x0C                       # corrupt a block header's integrity bits and fix nothing else
    assert prod.decode(bytes(bad))[0] == 1
    st, info = prod.blob_info(blob)
    info2 = dict(info); info2["nMasks"] = 0
    assert prod.decode(blob, info=info2)[0] == 0


def test_device_pointers(libs):
    """the same C API with CUDA device pointers for data, mask, blob and output"""
    import ctypes as C
    import torch
    prod, orc = libs
    img = c2_raster(300, 517)
    s_o, b_o, _ = orc.encode(img, 0.01)
    d_img = torch.from_numpy(img).cuda()
    d_out = torch.full((img.nbytes,), 0xAB, dtype=torch.uint8, device="cuda")
    n = C.c_uint(0)
    st = prod.f["encode"](d_img.data_ptr(), 6, 1, 517, 300, 1, 0, None, 0.01, d_out.data_ptr(), d_out.numel(), C.addressof(n))
    assert st == 0 and n.value == len(b_o)
    host = d_out.cpu().numpy()
    assert host[: n.value].tobytes() == b_o and not host[n.value:].any()
    d_dec = torch.empty((300, 517), dtype=torch.float32, device="cuda")
    d_mask = torch.empty((300, 517), dtype=torch.uint8, device="cuda")
    st = prod.f["decode"](d_out.data_ptr(), n.value, 1, d_mask.data_ptr(), 1, 517, 300, 1, 6, d_dec.data_ptr())
    assert st == 0
    _, d_ref, _ = orc.decode(b_o)
    assert np.array_equal(d_dec.cpu().numpy().view(np.uint8), d_ref[0, :, :, 0].view(np.uint8))
    assert bool((d_mask == 1).all())
    info = np.zeros(11, np.uint32)
    assert prod.f["getBlobInfo"](d_out.data_ptr(), n.value, info.ctypes.data, None, 11, 0) == 0 and info[3] == 517


def test_kernels_actually_ran(libs):
    import sys
    sys.path.insert(0, ROOT)
    import lerc_b200
    before = lerc_b200.stats()
    prod, _ = libs
    st, blob, _ = prod.encode(c2_raster(64, 64), 0.01)
    assert st == 0 and prod.decode(blob)[0] == 0
    after = lerc_b200.stats()
    assert after[0] >= before[0] + 2 and after[1] == before[1] + 1 and after[2] == before[2] + 1


def test_against_reference_library_if_present(libs):
    ref = ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libLerc_ref.so absent")
    prod, _ = libs
    for name, arr, mz, kw in CASES[:12]:
        s_r, b_r, _ = ref.encode(arr, mz, **kw)
        s_p, b_p, _ = prod.encode(arr, mz, **kw)
        assert s_r == s_p and b_r == b_p, name
        _, d_r, m_r = ref.decode(b_r)
        _, d_p, m_p = prod.decode(b_r)
        assert np.array_equal(d_r.view(np.uint8), d_p.view(np.uint8)), name


def test_medium_c2_and_c4(libs):
    prod, orc = libs
    img = c2_raster(1024, 1024)
    s_o, b_o, _ = orc.encode(img, 0.01)
    s_p, b_p, _ = prod.encode(img, 0.01)
    assert s_p == 0 and b_p == b_o, _first_diff(b_p, b_o)
    _, d_o, _ = orc.decode(b_o)
    _, d_p, _ = prod.decode(b_o)
    assert np.array_equal(d_o.view(np.uint8), d_p.view(np.uint8))
    rgb = c4_raster(512, 640)
    s_o, b_o, _ = orc.encode(rgb, 0, n_depth=3)
    s_p, b_p, _ = prod.encode(rgb, 0, n_depth=3)
    assert s_p == 0 and b_p == b_o, _first_diff(b_p, b_o)
    _, d_p, _ = prod.decode(b_o)
    assert np.array_equal(d_p[0], rgb)
