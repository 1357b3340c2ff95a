"""GPU tests (-m gpu) of the masked-raster paths added in round 2: the parallel byte RLE of the mask (encode: run / span boundaries by
prefix scans; decode: token list + parallel fill; RLE.cpp:32-331) and the block offsets of masked bands from the stream decoder's
boundary discovery + block-parallel verification (lerc_decode_stream.cuh, k_verify_offsets).  Blob bytes, decoded pixels and masks
must equal the oracle's; the coast-shaped cases must not end in the serial walk."""
import os
import struct

import numpy as np
import pytest

from cases import c2_raster
from lercapi import ROOT, fletcher32, oracle_lib, product_lib

pytestmark = pytest.mark.gpu
H, W = 300, 1024                       # 307200 pixels -> 38400 mask bytes (above the 4 KB where the parallel RLE takes over)


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib()
    assert prod is not None and orc is not None
    return prod, orc


def _mask_from_bytes(b):
    """MSB-first bit mask bytes -> byte mask [H][W]"""
    return np.unpackbits(b.astype(np.uint8))[: H * W].reshape(H, W).astype(np.uint8)


def _rle_cases():
    rng = np.random.default_rng(9)
    n_b = H * W // 8
    out = []
    m = np.ones((H, W), np.uint8); m[50:200, 100:900] = 0
    out.append(("rect_hole", m))
    out.append(("noise", (rng.random((H, W)) < 0.5).astype(np.uint8)))
    out.append(("sparse_invalid", (rng.random((H, W)) < 0.999).astype(np.uint8)))
    b = rng.integers(0, 256, n_b).astype(np.uint8)
    for k, run in enumerate([4, 5, 6, 7, 3]):                      # runs of exactly 4 / 5 / 6 bytes inside noise (RLE.cpp:74-79: 5 is the threshold)
        b[1000 + 50 * k: 1000 + 50 * k + run] = 0xAA
    for tail in (5, 6, 7):                                         # a run must start more than 5 bytes before the end to be a repeat
        bb = b.copy(); bb[-tail:] = 0x11
        out.append((f"runs_4_5_6_tail{tail}", _mask_from_bytes(bb)))
    b4 = rng.integers(0, 256, n_b).astype(np.uint8); b4[100:100 + 33000] = 0xF0
    out.append(("repeat_run_over_32767", _mask_from_bytes(b4)))     # pieces of at most 32767 (RLE.cpp:98-107)
    b5 = rng.integers(1, 255, n_b).astype(np.uint8)
    b5[1:] = np.where(b5[1:] == b5[:-1], b5[1:] ^ 1, b5[1:])
    out.append(("literal_span_over_32767", _mask_from_bytes(b5)))
    b6 = np.zeros(n_b, np.uint8); b6[::2] = 0xFF
    out.append(("alternating_bytes", _mask_from_bytes(b6)))
    b7 = np.repeat(rng.integers(0, 256, n_b // 5 + 1).astype(np.uint8), 5)[:n_b]
    out.append(("all_runs_of_5", _mask_from_bytes(b7)))
    b8 = np.full(n_b, 0xFF, np.uint8); b8[7] = 0x7F; b8[-1] = 0xFE
    out.append(("almost_all_valid", _mask_from_bytes(b8)))
    return out


RLE_CASES = _rle_cases()


@pytest.mark.parametrize("case", RLE_CASES, ids=[c[0] for c in RLE_CASES])
def test_mask_rle_and_masked_decode_match_oracle(libs, case):
    prod, orc = libs
    name, mask = case
    img = (np.random.default_rng(3).random((H, W)) * 100).astype(np.float32)
    s_o, b_o, _ = orc.encode(img, 0.01, mask=mask)
    s_p, b_p, _ = prod.encode(img, 0.01, mask=mask)
    assert s_o == 0 and s_p == 0 and b_p == b_o
    t_o, d_o, m_o = orc.decode(b_o)
    t_p, d_p, m_p = prod.decode(b_o)
    assert t_o == 0 and t_p == 0
    assert np.array_equal(m_p, m_o) and np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8))


def _coast(h, w, dtype=np.float32):
    img = c2_raster(h, w).astype(dtype)
    yy, xx = np.mgrid[0:h, 0:w]
    coast = w * 0.34 + w * 0.12 * np.sin(yy / 300.0) + 120 * np.sin(yy / 37.0) + 25 * np.sin(yy / 5.0)
    return img, (xx > coast).astype(np.uint8)


@pytest.mark.parametrize("shape,dtype,mz", [((1024, 2048), np.float32, 0.01), ((512, 4096), np.float32, 0.5), ((700, 1500), np.float64, 0.001),
                                            ((600, 2048), np.int16, 0), ((520, 1030), np.int32, 2)])
def test_coast_shaped_mask_takes_the_parallel_offsets(libs, shape, dtype, mz):
    """a ragged coast: hundreds of one-byte units (empty blocks) per block row and raw units of blocks with one or two valid pixels"""
    import sys
    sys.path.insert(0, ROOT)
    import lerc_b200
    prod, orc = libs
    img, mask = _coast(*shape, dtype=np.float32)
    if np.issubdtype(dtype, np.integer):
        img = np.clip(img * 3, np.iinfo(dtype).min, np.iinfo(dtype).max).astype(dtype)
    else:
        img = img.astype(dtype)
    s_o, b_o, _ = orc.encode(img, mz, mask=mask)
    s_p, b_p, _ = prod.encode(img, mz, mask=mask)
    assert s_o == 0 and s_p == 0 and b_p == b_o
    before = lerc_b200.stats()
    t_p, d_p, m_p = prod.decode(b_o)
    after = lerc_b200.stats()
    t_o, d_o, m_o = orc.decode(b_o)
    assert t_p == 0 and t_o == 0
    assert np.array_equal(m_p, m_o) and np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8))
    if not os.environ.get("LERC_B200_SIM"):           # (the counters are the real library's)
        assert after[4] == before[4] + 1, "block offsets came from the serial walk"


def test_corrupted_masked_blobs_fail_like_the_oracle(libs):
    prod, orc = libs
    img, mask = _coast(256, 2048)
    s_o, blob, _ = orc.encode(img, 0.01, mask=mask)
    assert s_o == 0
    rng = np.random.default_rng(17)
    for trial in range(40):
        bad = bytearray(blob)
        for _ in range(int(rng.integers(1, 4))):
            at = int(rng.integers(120, len(bad)))
            bad[at] ^= int(rng.integers(1, 256))
        struct.pack_into("<I", bad, 10, fletcher32(bytes(bad[14:])))       # checksum repaired so that the parsers are reached
        fixed = bytes(bad)
        t_o, d_o, m_o = orc.decode(fixed)
        t_p, d_p, m_p = prod.decode(fixed)
        assert t_p == t_o, f"trial {trial}: status {t_p} vs oracle {t_o}"
        if t_o == 0:
            assert np.array_equal(m_p, m_o) and np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8))
