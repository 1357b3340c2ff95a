"""GPU tests (-m gpu): lerc_encodeForVersion / lerc_computeCompressedSizeForVersion with codec versions 2..5
(Lerc::EncodeInternal_v5, Lerc.cpp:526-624; Lerc2::SetEncoderToOldVersion, Lerc2.cpp:52-63) and decoding of what they write,
version 2 included (MSB-first bit stuffing, no checksum: BitStuffer2.cpp:292-425).  The oracle is pinned to the reference for
these versions by tests/test_oracle_vs_reference.py::test_old_codec_versions_hashes."""
import os

import numpy as np
import pytest

from cases import all_cases, c4_raster
from lercapi import ROOT, oracle_lib, product_lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = all_cases(97, 131)


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib()
    assert prod is not None and orc is not None
    return prod, orc


@pytest.mark.parametrize("version", [2, 3, 4, 5])
def test_old_versions_match_oracle(libs, version):
    prod, orc = libs
    for name, arr, mz, kw in CASES:
        s_o, b_o, _ = orc.encode(arr, mz, version=version, **kw)
        s_p, b_p, buf = prod.encode(arr, mz, version=version, **kw)
        assert s_p == s_o, f"{name}: status {s_p} vs oracle {s_o}"
        if s_o != 0:
            continue
        assert b_p == b_o, f"{name}: blob differs ({len(b_p)} vs {len(b_o)} bytes)"
        assert not buf[len(b_p):].any(), name
        st, n = prod.compute_size(arr, mz, version=version, **kw)
        assert st == 0 and n == len(b_o), name
        t_o, d_o, m_o = orc.decode(b_o)
        t_p, d_p, m_p = prod.decode(b_o)
        assert t_o == 0 and t_p == 0, name
        assert np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8)), f"{name}: decoded pixels differ"
        assert (m_o is None) == (m_p is None) and (m_o is None or np.array_equal(m_o, m_p)), name


def test_version_argument_rules(libs):
    prod, orc = libs
    img = CASES[0][1]
    for v in (0, 1, 7, 99):                                            # Lerc2.cpp:54
        assert prod.encode(img, 0.01, version=v)[0] == orc.encode(img, 0.01, version=v)[0] == 2
    for v in (-1, -7, 6):                                              # any negative value and 6: the current codec
        s_p, b_p, _ = prod.encode(img, 0.01, version=v)
        s_o, b_o, _ = orc.encode(img, 0.01)
        assert s_p == 0 and b_p == b_o
    rgb = c4_raster(40, 50)
    for v in (2, 3):                                                   # nDepth > 1 needs codec version >= 4 (Lerc2.cpp:85-86)
        assert prod.encode(rgb, 0, n_depth=3, version=v)[0] == orc.encode(rgb, 0, n_depth=3, version=v)[0] == 1


@pytest.mark.parametrize("version", [4, 5])
def test_partial_nan_pixels_old_versions(libs, version):
    """nDepth > 1 with NaN in some depths of a pixel: version 6 reports ErrCode::NaN, versions <= 5 replace the NaN by -FLT_MAX
    (Lerc::ReplaceNaNValues, Lerc.cpp:901-939)"""
    prod, orc = libs
    rng = np.random.default_rng(3)
    a = (rng.random((60, 70, 3)) * 100).astype(np.float32)
    a[5:9, 10:30, 1] = np.nan
    a[20:22, 3:8, :] = np.nan
    assert prod.encode(a, 0.01, n_depth=3)[0] == orc.encode(a, 0.01, n_depth=3)[0] == 4
    s_o, b_o, _ = orc.encode(a, 0.01, n_depth=3, version=version)
    s_p, b_p, _ = prod.encode(a, 0.01, n_depth=3, version=version)
    assert s_o == 0 and s_p == 0 and b_p == b_o
    t_o, d_o, m_o = orc.decode(b_o)
    t_p, d_p, m_p = prod.decode(b_o)
    assert t_o == 0 and t_p == 0 and np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8)) and np.array_equal(m_o, m_p)


def test_bluemarble_reencode_v3_reproduces_the_shipped_blob(libs):
    """encodeForVersion(3) of the pixels decoded from testData/bluemarble_256_256_3_byte.lerc2 reproduces the file
    except, per band, the checksum and the 4-byte Huffman read-ahead pad."""
    prod, orc = libs
    blob = open(os.path.join(GOLD, "bluemarble_256_256_3_byte.lerc2"), "rb").read()
    st, data, mask = prod.decode(blob)
    assert st == 0
    n_bands = data.shape[0]
    st, mine, _ = prod.encode(np.ascontiguousarray(data[:, :, :, 0]), 0, n_bands=n_bands, mask=mask, version=3)
    assert st == 0 and len(mine) == len(blob)
    a, b = np.frombuffer(mine, np.uint8).copy(), np.frombuffer(blob, np.uint8).copy()
    pos = 0
    for _ in range(n_bands):
        size = int(np.frombuffer(blob[pos + 30:pos + 34], np.int32)[0])          # v3 header: blobSize at byte 30
        for buf in (a, b):
            buf[pos + 10:pos + 14] = 0
            buf[pos + size - 4:pos + size] = 0
        pos += size
    assert pos == len(blob) and np.array_equal(a, b)
