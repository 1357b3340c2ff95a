"""GPU tests (-m gpu) of the tile batch entry points lerc_b200_encodeTiles / lerc_b200_decodeTiles (include/lerc_b200.h).

Contract: blob t of the batch == lerc_encode of tile t's pixel window alone (checked against the oracle, byte for byte),
and the batch decode == the oracle's lerc_decode of every blob (bit-exact).  The cases include tiles that flip the
encoder's image-global decisions (constant, all-integer, pre-rounded, NaN, LUT, 16x16, one sweep), edge tiles, all pixel
types, and BASELINE config 5's tile shape (256 x 256 float32, maxZError 0.01) at a size the scalar oracle samples."""
import ctypes as C

import numpy as np
import pytest

from cases import c2_raster, tile_cases
from lercapi import DT_CODE, decode_tiles, encode_tiles, oracle_lib, product_lib, tile_windows, tiles_api

pytestmark = pytest.mark.gpu
CASES = tile_cases()


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib()
    assert prod is not None, "lerc_b200/libLerc.so.4 missing"
    assert orc is not None, "oracle/_build/liblerc_oracle.so missing"
    return prod, orc


def _oracle_blobs(orc, raster, tr, tc, mz):
    blobs = []
    for ys, xs in tile_windows(raster.shape[0], raster.shape[1], tr, tc):
        st, b, _ = orc.encode(np.ascontiguousarray(raster[ys, xs]), mz)
        assert st == 0
        blobs.append(b)
    return blobs


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_encode_tiles_matches_oracle(libs, case):
    prod, orc = libs
    name, raster, tr, tc, mz = case
    st, blobs, off = encode_tiles(prod, raster, tr, tc, mz)
    assert st == 0
    want = _oracle_blobs(orc, raster, tr, tc, mz)
    assert len(blobs) == len(want)
    for t, (a, b) in enumerate(zip(blobs, want)):
        assert a == b, f"tile {t}: {len(a)} vs {len(b)} bytes"


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_decode_tiles_matches_oracle(libs, case):
    prod, orc = libs
    name, raster, tr, tc, mz = case
    want = _oracle_blobs(orc, raster, tr, tc, mz)
    st, dec = decode_tiles(prod, want, raster.dtype, raster.shape[0], raster.shape[1], tr, tc)
    assert st == 0
    for t, (ys, xs) in enumerate(tile_windows(raster.shape[0], raster.shape[1], tr, tc)):
        st_o, d_o, _ = orc.decode(want[t])
        assert st_o == 0
        assert np.array_equal(dec[ys, xs].view(np.uint8), d_o[0, :, :, 0].view(np.uint8)), f"tile {t}"


def test_buffer_too_small_and_bad_params(libs):
    prod, _ = libs
    api = tiles_api(prod)
    img = c2_raster(128, 128)
    st, blobs, off = encode_tiles(prod, img, 64, 64, 0.01)
    assert st == 0
    st2, _, _ = encode_tiles(prod, img, 64, 64, 0.01, buf_size=int(off[-1]) - 1)
    assert st2 == 3                                                     # BufferTooSmall
    st3, b3, off3 = encode_tiles(prod, img, 64, 64, 0.01, buf_size=int(off[-1]))
    assert st3 == 0 and b3 == blobs
    n = C.c_ulonglong(0)
    o = np.zeros(5, np.uint64)
    out = np.zeros(1024, np.uint8)
    assert api.lerc_b200_encodeTiles(None, 6, 128, 128, 64, 64, 0.01, out.ctypes.data, 1024, o.ctypes.data, C.addressof(n)) == 2
    assert api.lerc_b200_encodeTiles(img.ctypes.data, 6, 128, 128, 0, 64, 0.01, out.ctypes.data, 1024, o.ctypes.data, C.addressof(n)) == 2
    assert api.lerc_b200_encodeTiles(img.ctypes.data, 6, 128, 128, 64, 64, -1.0, out.ctypes.data, 1024, o.ctypes.data, C.addressof(n)) == 2


def test_corrupted_tile_fails_like_lerc_decode(libs):
    prod, orc = libs
    img = c2_raster(128, 192)
    blobs = _oracle_blobs(orc, img, 64, 64, 0.01)
    bad = bytearray(blobs[3])
    bad[200] ^= 0x55                                                     # checksum no longer matches
    st_o, _, _ = orc.decode(bytes(bad))
    assert st_o != 0
    st, _ = decode_tiles(prod, blobs[:3] + [bytes(bad)] + blobs[4:], np.float32, 128, 192, 64, 64)
    assert st == st_o


def test_corrupted_constant_tile_fails_like_lerc_decode(libs):
    """a header-only blob (constant tile) with one byte changed: the batch decoder and the oracle agree on the verdict for every byte
    position tried, and on the pixels where the change is harmless"""
    prod, orc = libs
    img = c2_raster(128, 192)
    img[0:64, 64:128] = 3.25
    blobs = _oracle_blobs(orc, img, 64, 64, 0.01)
    assert len(blobs[1]) == 94
    rng = np.random.default_rng(3)
    for _ in range(24):
        bad = bytearray(blobs[1])
        k = int(rng.integers(0, 94))
        bad[k] ^= int(rng.integers(1, 256))
        st_o, d_o, _ = orc.decode(bytes(bad))
        st, dec = decode_tiles(prod, [blobs[0], bytes(bad)] + blobs[2:], np.float32, 128, 192, 64, 64)
        assert (st == 0) == (st_o == 0), f"byte {k}: status {st} vs oracle {st_o}"
        if st == 0:
            assert np.array_equal(dec[0:64, 64:128].view(np.uint8), d_o[0, :, :, 0].view(np.uint8))


def test_device_pointers_config5_shape(libs):
    """BASELINE config 5's tile shape on device-resident buffers: 2048 x 4096 float32 as 128 tiles of 256 x 256."""
    import torch
    import lerc_b200
    prod, orc = libs
    api = tiles_api(prod)
    h, w, t = 2048, 4096, 256
    img = c2_raster(h, w)
    n_tiles = (h // t) * (w // t)
    d_img = torch.from_numpy(img).cuda()
    cap = int(api.lerc_b200_tilesMaxBytes(6, w, h, t, t))
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(n_tiles + 1, dtype=torch.int64, device="cuda")
    n = C.c_ulonglong(0)
    torch.cuda.synchronize()
    s0 = lerc_b200.stats()
    st = api.lerc_b200_encodeTiles(d_img.data_ptr(), 6, w, h, t, t, 0.01, d_out.data_ptr(), cap, d_off.data_ptr(), C.addressof(n))
    s1 = lerc_b200.stats()
    assert st == 0
    assert s1[3] - s0[3] == n_tiles, "the fused batch pass did not code every tile"
    assert s1[0] - s0[0] <= 4, "more launches than the batch path needs"
    off = d_off.cpu().numpy()
    out = d_out[: n.value].cpu().numpy()
    assert off[0] == 0 and off[-1] == n.value
    wins = list(tile_windows(h, w, t, t))
    for k in (0, 1, 17, 63, n_tiles - 1):                               # the scalar oracle samples the tiles
        ys, xs = wins[k]
        st_o, b_o, _ = orc.encode(np.ascontiguousarray(img[ys, xs]), 0.01)
        assert st_o == 0 and out[off[k]:off[k + 1]].tobytes() == b_o, f"tile {k}"
    # every blob is accepted by the product's own single-image decoder path, and the batch decode agrees with it
    d_dec = torch.full((h, w), -1.0, dtype=torch.float32, device="cuda")
    st = api.lerc_b200_decodeTiles(d_out.data_ptr(), n.value, d_off.data_ptr(), 6, w, h, t, t, d_dec.data_ptr())
    s2 = lerc_b200.stats()
    assert st == 0 and s2[4] - s1[4] == n_tiles
    dec = d_dec.cpu().numpy()
    assert float(np.abs(dec.astype(np.float64) - img).max()) <= 0.01 * 1.1          # float32 rounding of the decoded value; the reference's own slack (Lerc.cpp:1137)
    for k in (0, 5, 64, n_tiles - 1):
        ys, xs = wins[k]
        st_o, d_o, _ = orc.decode(out[off[k]:off[k + 1]].tobytes())
        assert st_o == 0 and np.array_equal(dec[ys, xs].view(np.uint8), d_o[0, :, :, 0].view(np.uint8)), f"tile {k}"


def test_corrupted_tiles_fuzz(libs):
    """1-3 flipped bytes in one tile's blob, checksum repaired so that the header parser, the walker and the block decoder of the batch
    path are reached: status and pixels of lerc_b200_decodeTiles equal the oracle's lerc_decode of that blob"""
    prod, orc = libs
    fl = orc.lib.lo_fletcher32
    fl.restype = C.c_uint32
    fl.argtypes = [C.c_void_p, C.c_int]
    img = c2_raster(128, 192)
    img[70:100, :] = 5.0
    wins = list(tile_windows(128, 192, 64, 64))
    blobs = _oracle_blobs(orc, img, 64, 64, 0.01)
    rng = np.random.default_rng(9)
    for it in range(60):
        t = int(rng.integers(0, len(blobs)))
        b = bytearray(blobs[t])
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(14, len(b)))] ^= int(rng.integers(1, 256))
        buf = np.frombuffer(bytes(b), np.uint8).copy()
        cs = fl(buf[14:].ctypes.data, len(b) - 14)
        buf[10:14] = np.frombuffer(np.uint32(cs).tobytes(), np.uint8)
        s_o, d_o, _ = orc.decode(buf.tobytes())
        s_p, d_p = decode_tiles(prod, blobs[:t] + [buf.tobytes()] + blobs[t + 1:], np.float32, 128, 192, 64, 64)
        assert (s_p == 0) == (s_o == 0), f"case {it}, tile {t}: status {s_p} vs oracle {s_o}"
        if s_o == 0:
            ys, xs = wins[t]
            assert np.array_equal(d_p[ys, xs].view(np.uint8), d_o[0, :, :, 0].view(np.uint8)), f"case {it}, tile {t}"
