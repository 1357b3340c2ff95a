"""GPU tests at BASELINE.json sizes: parity against the reference where the CPU side finishes in seconds, and
size-independent properties (round-trip error bound, idempotence of decode(encode(decode(.))), size query == bytes
written, checksum acceptance by the oracle) where it does not."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from cases import c2_raster, c4_raster
from lercapi import ROOT, oracle_lib, product_lib, ref_lib

sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib()
    assert prod is not None and orc is not None
    return prod, orc


def test_c2_full_size_bit_exact(libs):
    """BASELINE configs[1]: 4096 x 4096 float32, maxZError 0.01 -- blob and pixels bit-exact vs the reference / oracle"""
    import lerc_b200
    prod, orc = libs
    chk = ref_lib() or orc
    img = c2_raster(4096, 4096)
    s_r, b_r, _ = chk.encode(img, 0.01)
    s0 = lerc_b200.stats()
    s_p, b_p, _ = prod.encode(img, 0.01)
    assert s_r == 0 and s_p == 0 and b_p == b_r
    st, n = prod.compute_size(img, 0.01)
    assert st == 0 and n == len(b_r)
    _, d_r, _ = chk.decode(b_r)
    st, d_p, _ = prod.decode(b_r)
    s1 = lerc_b200.stats()
    assert st == 0 and np.array_equal(d_p.view(np.uint8), d_r.view(np.uint8))
    assert s1[3] == s0[3] + 1 and s1[4] == s0[4] + 1, "fused paths were not taken"
    assert float(np.abs(d_p[0, :, :, 0].astype(np.float64) - img).max()) <= 0.01 * 1.1     # reference slack (Lerc.cpp:1137)


def test_c4_full_size_bit_exact(libs):
    """BASELINE configs[3]: 8192 x 8192 uint8 RGB (nDepth = 3), lossless -- the 8-bit Huffman path: blob byte-exact vs the
    reference library (oracle/_ref, else the oracle), decode of the reference's blob bit-exact and lossless"""
    prod, orc = libs
    chk = ref_lib() or orc
    img = c4_raster(8192, 8192)
    s_r, b_r, _ = chk.encode(img, 0, n_depth=3)
    s_p, b_p, _ = prod.encode(img, 0, n_depth=3)
    assert s_r == 0 and s_p == 0
    assert len(b_p) == len(b_r) and b_p == b_r
    mode_at = 90 + 4 + 2 * 3 + 1                                  # header | mask byte count | ranges | one-sweep flag
    assert b_r[mode_at] in (1, 2), "expected a Huffman image mode for this raster"
    st, n = prod.compute_size(img, 0, n_depth=3)
    assert st == 0 and n == len(b_r)
    st, d_p, _ = prod.decode(b_r)
    assert st == 0 and np.array_equal(d_p[0], img)
    bad = bytearray(b_r); bad[len(bad) // 2] ^= 0x10
    assert prod.decode(bytes(bad))[0] != 0


def test_c3_band_properties_device_resident(libs):
    """BASELINE configs[2] shape, one 16384 x 16384 float32 band (1 GiB) at maxZError 0.001, device resident:
    error bound, decode(encode(decode(x))) == decode(x) (idempotence), a strip of the blob's pixels equals the oracle's."""
    import torch
    prod, orc = libs
    n = 16384
    torch.manual_seed(5)
    xx = torch.arange(n, device="cuda", dtype=torch.float32)
    img = (1000 + 300 * torch.sin(xx[None, :] / 97) * torch.cos(xx[:, None] / 131) + 50 * torch.sin(xx[None, :] / 13 + xx[:, None] / 17)
           + 0.5 * torch.randn(n, n, device="cuda")).contiguous()
    # the library works on its own (non-blocking) stream unless lerc_b200_set_stream is used: the raster must be complete
    torch.cuda.synchronize()
    cap = n * n * 4 + (1 << 20)
    blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
    out = torch.empty_like(img)
    nb = C.c_uint(0)
    enc, dec = prod.f["encode"], prod.f["decode"]
    assert enc(img.data_ptr(), 6, 1, n, n, 1, 0, None, 0.001, blob.data_ptr(), cap, C.addressof(nb)) == 0
    assert dec(blob.data_ptr(), nb.value, 0, None, 1, n, n, 1, 6, out.data_ptr()) == 0
    assert float((out.double() - img.double()).abs().max().item()) <= 0.001 * 1.1
    blob2 = torch.empty(cap, dtype=torch.uint8, device="cuda")
    out2 = torch.empty_like(img)
    nb2 = C.c_uint(0)
    assert enc(out.data_ptr(), 6, 1, n, n, 1, 0, None, 0.001, blob2.data_ptr(), cap, C.addressof(nb2)) == 0
    assert dec(blob2.data_ptr(), nb2.value, 0, None, 1, n, n, 1, 6, out2.data_ptr()) == 0
    assert float((out2.double() - out.double()).abs().max().item()) <= 0.001 * 1.1
    # a 64-row strip re-encoded alone: bit-exact vs the oracle, pixels too
    strip = img[4096:4160].contiguous().cpu().numpy()
    s_o, b_o, _ = orc.encode(strip, 0.001)
    s_p, b_p, _ = prod.encode(strip, 0.001)
    assert s_o == 0 and s_p == 0 and b_o == b_p
    info = np.zeros(11, np.uint32)
    assert prod.f["getBlobInfo"](blob.data_ptr(), nb.value, info.ctypes.data, None, 11, 0) == 0
    assert info[3] == n and info[4] == n and info[7] == nb.value


def test_four_bands_device_resident(libs):
    """nBands = 4 (configs[2] layout) at a size the oracle handles: one call, 4 band blobs back to back"""
    prod, orc = libs
    bands = np.stack([c2_raster(1024, 1024, seed=1234 + b, phase=0.1 * b) for b in range(4)])
    s_o, b_o, _ = orc.encode(bands, 0.001, n_bands=4)
    s_p, b_p, _ = prod.encode(bands, 0.001, n_bands=4)
    assert s_o == 0 and s_p == 0 and b_o == b_p
    st, d_p, _ = prod.decode(b_o)
    _, d_o, _ = orc.decode(b_o)
    assert st == 0 and np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8))
