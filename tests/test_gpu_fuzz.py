"""GPU robustness test: corrupted blobs (with the checksum repaired so that the parsers are reached) must be rejected
or decoded exactly like the oracle does -- never a crash, hang or different pixels."""
import struct

import numpy as np
import pytest

from cases import c2_raster, c4_raster
from lercapi import oracle_lib, product_lib

pytestmark = pytest.mark.gpu


def fletcher32(data):
    """Lerc2.cpp:1037-1064 (big-endian 16-bit words, sums start at 0xffff, odd tail byte << 8)"""
    s1 = s2 = 0xffff
    n = len(data) // 2
    i = 0
    while n:
        t = min(359, n)
        n -= t
        for _ in range(t):
            s1 += (data[i] << 8) + data[i + 1]
            s2 += s1
            i += 2
        s1 = (s1 & 0xffff) + (s1 >> 16)
        s2 = (s2 & 0xffff) + (s2 >> 16)
    if len(data) & 1:
        s1 += data[i] << 8
        s2 += s1
    s1 = (s1 & 0xffff) + (s1 >> 16)
    s2 = (s2 & 0xffff) + (s2 >> 16)
    return ((s2 << 16) | s1) & 0xffffffff


def repair(blob):
    b = bytearray(blob)
    size = struct.unpack_from("<i", b, 34)[0]          # v6 header: blobSize at byte 34 (6 + 4 + 4 + 5 * 4)
    if 14 < size <= len(b):
        struct.pack_into("<I", b, 10, fletcher32(bytes(b[14:size])))
    return bytes(b)


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib()
    assert prod is not None and orc is not None
    return prod, orc


def _cases():
    rng = np.random.default_rng(11)
    out = [("f32", c2_raster(264, 520), 0.01, {}),
           ("i16", np.clip(c2_raster(200, 333) * 3 - 2000, -32768, 32767).astype(np.int16), 0, {}),
           ("f32_masked", c2_raster(128, 256), 0.01, {"mask": (rng.random((128, 256)) > 0.1).astype(np.uint8)}),
           ("u8x3", c4_raster(96, 128), 0, {"n_depth": 3}),
           ("u8x3_masked", c4_raster(64, 96), 0, {"n_depth": 3, "mask": (rng.random((64, 96)) > 0.2).astype(np.uint8)}),
           ("f64", c2_raster(72, 200).astype(np.float64), 0.001, {}),
           ("f32_2bands", np.stack([c2_raster(64, 128, seed=1), c2_raster(64, 128, seed=2)]), 0.01, {"n_bands": 2})]
    return out


def test_repaired_checksum_helper_matches(libs):
    prod, orc = libs
    st, blob, _ = orc.encode(c2_raster(40, 40), 0.01)
    assert repair(blob) == blob


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_corrupted_streams_agree_with_oracle(libs, case):
    prod, orc = libs
    name, arr, mz, kw = case
    st, blob, _ = orc.encode(arr, mz, **kw)
    assert st == 0
    st, info = orc.blob_info(blob)
    rng = np.random.default_rng(5)
    n_ok = n_fail = 0
    for trial in range(60):
        b = bytearray(blob)
        k = int(rng.integers(1, 4))
        for _ in range(k):
            pos = int(rng.integers(95, len(b)))          # past the header: the stream, mask and ranges
            b[pos] ^= int(rng.integers(1, 256))
        bad = repair(bytes(b))
        s_o, d_o, m_o = orc.decode(bad, info=info)
        s_p, d_p, m_p = prod.decode(bad, info=info)
        assert (s_p == 0) == (s_o == 0), f"{name} trial {trial}: status {s_p} vs oracle {s_o}"
        if s_o == 0:
            n_ok += 1
            assert np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8)), f"{name} trial {trial}: pixels differ"
        else:
            n_fail += 1
    assert n_ok + n_fail == 60
