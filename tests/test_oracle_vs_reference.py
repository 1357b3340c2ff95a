"""CPU tests (-m "not gpu"): pin the plain-C oracle (oracle/lerc_oracle.c) to the reference.

 (a) golden fixtures: the reference's shipped blobs + the JS sanity blob decode to exactly what the
     reference decoded (tests/golden/*.npz, made by tests/golden/make_golden.py);
 (b) synthetic rasters: the oracle's blob and decoded pixels hash to the reference's (synthetic_ref.npz);
 (c) when oracle/_ref/libLerc_ref.so is present (always in the build container), byte-for-byte against it.
"""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from cases import all_cases, bitplane_cases, fpl_cases, fpl_encode_cases, fpl_fuzz_cases, nodata_cases
from lercapi import ROOT, oracle_lib, ref_lib

GOLD = os.path.join(ROOT, "tests", "golden")
FPL_CASES = {"f32_lossless_raw", "f64_lossless_raw"}   # reference uses the FPL codec here: sizes differ by design


@pytest.fixture(scope="module")
def oracle():
    if oracle_lib() is None:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = oracle_lib(fpl_encoder=True)
    assert lib is not None
    return lib


@pytest.mark.parametrize("name", ["california_400_400_1_float", "bluemarble_256_256_3_byte", "js_sanity_30_20_3_byte"])
def test_golden_decode(oracle, name):
    blob = open(os.path.join(GOLD, name + ".lerc2"), "rb").read()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    st, info = oracle.blob_info(blob)
    assert st == 0
    want = dict(zip([str(k) for k in g["info_keys"]], g["info"]))
    for k, v in want.items():
        assert float(info[k]) == float(v), (k, info[k], v)
    st, data, mask = oracle.decode(blob)
    assert st == 0
    assert np.array_equal(data.view(np.uint8), g["data"].view(np.uint8))      # bit-exact, floats included
    if g["mask"].size:
        assert np.array_equal(mask, g["mask"])
    st, mins, maxs = oracle.data_ranges(blob, info["nDepth"], info["nBands"])
    assert st == 0 and np.array_equal(mins, g["mins"]) and np.array_equal(maxs, g["maxs"])


def test_js_sanity_expectations(oracle):
    """the assertions of OtherLanguages/js/tests/sanity.mjs:27-40"""
    blob = open(os.path.join(GOLD, "js_sanity_30_20_3_byte.lerc2"), "rb").read()
    st, info = oracle.blob_info(blob)
    assert (info["nCols"], info["nRows"], info["nDepth"], info["dataType"]) == (30, 20, 3, 1)
    st, data, _ = oracle.decode(blob)
    assert data.reshape(-1)[:6].tolist() == [13, 57, 68, 14, 59, 80]
    px = data[0]
    assert [int(px[..., d].min()) for d in range(3)] == [0, 30, 60] and int(px.max()) == 89


def test_stored_checksums(oracle):
    """free known-answer test: the Fletcher-32 values stored in the shipped blobs"""
    import ctypes
    f = oracle.lib.lo_fletcher32
    f.restype = ctypes.c_uint32
    f.argtypes = [ctypes.c_void_p, ctypes.c_int]
    for name, size in [("california_400_400_1_float", 176451), ("bluemarble_256_256_3_byte", 18747)]:
        blob = np.frombuffer(open(os.path.join(GOLD, name + ".lerc2"), "rb").read(), np.uint8)
        stored = int(np.frombuffer(blob[10:14].tobytes(), np.uint32)[0])
        assert f(blob[14:].ctypes.data, size - 14) == stored


def test_synthetic_hashes(oracle):
    g = np.load(os.path.join(GOLD, "synthetic_ref.npz"))
    want = {str(n): (int(s), str(e), str(d)) for n, s, e, d in zip(g["names"], g["sizes"], g["enc"], g["dec"])}
    checked = 0
    for name, arr, mz, kw in all_cases():
        if name not in want:
            continue
        st, blob, buf = oracle.encode(arr, mz, **kw)
        assert st == 0, name
        assert not buf[len(blob):].any(), f"{name}: output buffer not zero-filled past the blob"
        st, data, mask = oracle.decode(blob)
        assert st == 0, name
        if name not in FPL_CASES:
            assert len(blob) == want[name][0], name
            assert hashlib.sha256(blob).hexdigest() == want[name][1], f"{name}: blob differs from the reference's"
            h = hashlib.sha256(data.tobytes())
            if mask is not None:
                h.update(mask.tobytes())
            assert h.hexdigest() == want[name][2], f"{name}: decoded pixels differ from the reference's"
        else:     # lossless float without FPL: must still round-trip exactly
            assert np.array_equal(data.reshape(arr.shape).view(np.uint8), np.ascontiguousarray(arr).view(np.uint8))
        st2, n = oracle.compute_size(arr, mz, **kw)
        assert st2 == 0 and n == len(blob), name
        checked += 1
    assert checked >= 50


def test_old_codec_versions_hashes(oracle):
    """lerc_encodeForVersion(2..5) (Lerc::EncodeInternal_v5, Lerc.cpp:526-624; v2 = MSB-first bit stuffing without checksum):
    status, blob and decoded pixels of the oracle hash to the reference's (synthetic_ref_versions.npz)"""
    g = np.load(os.path.join(GOLD, "synthetic_ref_versions.npz"))
    want = {str(k): (int(st), int(s), str(e), str(d)) for k, st, s, e, d in zip(g["keys"], g["status"], g["sizes"], g["enc"], g["dec"])}
    checked = 0
    for v in (2, 3, 4, 5):
        for name, arr, mz, kw in all_cases():
            st_w, size_w, enc_w, dec_w = want[f"{name}|{v}"]
            st, blob, _ = oracle.encode(arr, mz, version=v, **kw)
            assert st == st_w, (name, v)
            if st != 0:
                continue
            assert len(blob) == size_w and hashlib.sha256(blob).hexdigest() == enc_w, f"{name} v{v}: blob differs from the reference's"
            st, data, mask = oracle.decode(blob)
            assert st == 0
            h = hashlib.sha256(data.tobytes())
            if mask is not None:
                h.update(mask.tobytes())
            assert h.hexdigest() == dec_w, f"{name} v{v}: decoded pixels differ from the reference's"
            st2, n = oracle.compute_size(arr, mz, version=v, **kw)
            assert st2 == 0 and n == len(blob)
            checked += 1
    assert checked >= 200


NODATA_FPL_CASES = {"f32_d3_nodata_far_lossless", "f32_d3_nodata_close"}      # the filter falls back to lossless float: FPL in the reference


def test_nodata_4d_hashes(oracle):
    """lerc_encode_4D / lerc_decode_4D with per-band noData values: status, blob, decoded pixels and the returned noData arrays of the
    oracle equal the reference's (nodata_ref.npz); where the reference switches to its FPL codec the oracle must still round-trip"""
    g = np.load(os.path.join(GOLD, "nodata_ref.npz"))
    want = {str(n): (int(st), int(s), str(e), str(d), str(u), str(v)) for n, st, s, e, d, u, v in
            zip(g["names"], g["status"], g["sizes"], g["enc"], g["dec"], g["uses"], g["vals"])}
    for name, arr, mz, kw in nodata_cases():
        st_w, size_w, enc_w, dec_w, uses_w, vals_w = want[name]
        st, blob = oracle.encode_4d(arr, mz, **kw)
        assert st == st_w, name
        if st != 0:
            continue
        st, data, mask, uses, vals = oracle.decode_4d(blob)
        assert st == 0, name
        assert uses.tobytes().hex() == uses_w and vals.tobytes().hex() == vals_w, name
        if name in NODATA_FPL_CASES:
            a = np.ascontiguousarray(arr).reshape(data.shape)
            valid = np.ones(a.shape[1:3], bool) if mask is None else mask[0].astype(bool)
            assert np.array_equal(data[0][valid].view(np.uint8), a[0][valid].view(np.uint8)), name
            continue
        assert len(blob) == size_w and hashlib.sha256(blob).hexdigest() == enc_w, f"{name}: blob differs from the reference's"
        h = hashlib.sha256(data.tobytes())
        if mask is not None:
            h.update(mask.tobytes())
        assert h.hexdigest() == dec_w, f"{name}: decoded pixels differ from the reference's"
        assert oracle.encode_4d(arr, mz, size_only=True, **kw) == (0, len(blob)), name
        info = oracle.blob_info(blob)[1]
        has = info["nUsesNoDataValue"] > 0 and info["nDepth"] > 1
        assert oracle.decode_4d(blob, want_no_data=False)[0] == (5 if has else 0), name      # ErrCode::HasNoData (Lerc.cpp:431-434)


def test_fpl_blobs_decode_like_the_reference(oracle):
    """blobs made by the reference's lossless float codec (IEM_DeltaDeltaHuffman, fpl_*.cpp): the oracle's decode hashes to the
    reference's (fpl_ref.npz), invalid pixels included, and equals the input on the valid pixels"""
    g = np.load(os.path.join(GOLD, "fpl_ref.npz"))
    n_fpl = 0
    for name, arr, kw in fpl_cases():
        blob = g["blob_" + name].tobytes()
        st, data, mask = oracle.decode(blob)
        assert st == 0, name
        h = hashlib.sha256(data.tobytes())
        if mask is not None:
            h.update(mask.tobytes())
        assert h.hexdigest() == str(g["hash_" + name]), f"{name}: decoded pixels differ from the reference's"
        n_fpl += b"\x03" in blob[90:140]           # (the image-mode byte 3 sits right behind header, mask count and ranges)
    assert n_fpl >= 10


def test_fpl_encoder_makes_the_references_blobs(oracle):
    """float rasters at maxZError 0: the oracle's lossless float ENCODER (predictor choice from entropy estimates of test blocks,
    per-plane byte-delta level, Huffman / PackBits / raw / one-value planes, the 10 % rule against raw tiling) writes the blobs
    the reference wrote (fpl_ref.npz) - up to the 4 uninitialised bytes the reference leaves behind each Huffman-coded plane"""
    from lercapi import fpl_normalize
    g = np.load(os.path.join(GOLD, "fpl_ref.npz"))
    n_fpl = 0
    for name, arr, kw in fpl_cases():
        want = g["blob_" + name].tobytes()
        st, blob, _ = oracle.encode(arr, 0.0, **kw)
        assert st == 0, name
        assert fpl_normalize(blob) == blob, name
        assert blob == fpl_normalize(want), f"{name}: blob differs from the reference's"
        assert oracle.compute_size(arr, 0.0, **kw) == (0, len(want)), name
        n_fpl += b"\x03" in blob[90:140]
    assert n_fpl >= 10


def test_fpl_encoder_vs_live_reference(oracle):
    """the extra encode cases and the fuzz set of the GPU tests, oracle against oracle/_ref (the unmodified reference compiled in
    place) where that library exists -- it does not travel in git, so this one skips without it"""
    from lercapi import fpl_normalize, ref_lib
    ref = ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libLerc_ref.so not built")
    for name, arr, kw in fpl_encode_cases() + fpl_fuzz_cases():
        s_r, b_r, _ = ref.encode(arr, 0.0, **kw)
        s_o, b_o, _ = oracle.encode(arr, 0.0, **kw)
        assert s_r == s_o, name
        if s_r == 0:
            assert fpl_normalize(b_r) == b_o, name


def test_bitplane_mode_hashes(oracle):
    """maxZErr == 777 (Lerc2.cpp:210-217 -> TryBitPlaneCompression :1071-1229): status, blob and the maxZError the reference ends up
    with (bitplane_ref.npz)"""
    g = np.load(os.path.join(GOLD, "bitplane_ref.npz"))
    want = {str(n): (int(st), int(s), str(e), float(m)) for n, st, s, e, m in zip(g["names"], g["status"], g["sizes"], g["enc"], g["maxzerr"])}
    raised = 0
    for name, arr, kw in bitplane_cases():
        st_w, size_w, enc_w, mz_w = want[name]
        st, blob, _ = oracle.encode(arr, 777, **kw)
        assert st == st_w, name
        if st != 0:
            continue
        assert len(blob) == size_w and hashlib.sha256(blob).hexdigest() == enc_w, f"{name}: blob differs from the reference's"
        assert oracle.blob_info(blob)[1]["maxZErrUsed"] == mz_w, name
        raised += mz_w > 0.5
    assert raised >= 6


def test_bluemarble_reencode_v3_reproduces_the_shipped_blob(oracle):
    """lerc_encodeForVersion(3) of the pixels decoded from testData/bluemarble_256_256_3_byte.lerc2 reproduces the
    file except, per band, the checksum (4 bytes at offset 10) and the Huffman read-ahead pad (last 4 bytes)."""
    blob = open(os.path.join(GOLD, "bluemarble_256_256_3_byte.lerc2"), "rb").read()
    g = np.load(os.path.join(GOLD, "bluemarble_256_256_3_byte.npz"))
    data, mask = g["data"], g["mask"]
    n_bands = data.shape[0]
    st, mine, _ = oracle.encode(np.ascontiguousarray(data[:, :, :, 0]), 0, n_bands=n_bands, mask=mask if mask.size else None, version=3)
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


def test_against_reference_library(oracle):
    ref = ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libLerc_ref.so not built (no /root/reference on this machine)")
    for name, arr, mz, kw in all_cases():
        s1, b1, _ = ref.encode(arr, mz, **kw)
        s2, b2, _ = oracle.encode(arr, mz, **kw)
        assert s1 == s2, name
        if s1 != 0:
            continue
        if name not in FPL_CASES:
            assert b1 == b2, name
        for blob in (b1, b2):
            if name in FPL_CASES and blob is b1:
                continue              # FPL blobs are outside the oracle's scope
            t1, d1, m1 = ref.decode(blob)
            t2, d2, m2 = oracle.decode(blob)
            assert t1 == 0 and t2 == 0, name
            assert np.array_equal(d1.view(np.uint8), d2.view(np.uint8)) and np.array_equal(m1, m2), name


def test_error_codes(oracle):
    a = np.zeros((4, 4), np.float32)
    assert oracle.encode(a, -1.0)[0] == 2                       # WrongParam: maxZErr < 0 (Lerc_c_api_impl.cpp:82)
    st, blob, _ = oracle.encode(a + 1, 0.01, buf_size=50)
    assert st == 3                                              # BufferTooSmall (Lerc.cpp:764)
    nan = np.full((4, 4, 2), 1.0, np.float32); nan[0, 0, 0] = np.nan
    assert oracle.encode(nan, 0.01, n_depth=2)[0] == 4          # NaN: mixed NaN at one pixel, no noData value (Lerc.cpp:1481)
    st, blob, _ = oracle.encode(a + 1, 0.01)
    bad = bytearray(blob); bad[-1] ^= 0xFF
    assert oracle.decode(bytes(bad))[0] == 1                    # checksum mismatch -> Failed (Lerc2.cpp:599)
    assert oracle.decode(blob[:-3])[0] == 1                     # truncated


def test_rle_against_reference_tokens(oracle):
    """RLE restated as a run tokenizer: round-trips and hits the documented corner cases"""
    import ctypes
    lib = oracle.lib
    lib.lo_rle_size.restype = ctypes.c_size_t
    lib.lo_rle_encode.restype = ctypes.c_size_t
    lib.lo_rle_size.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.lo_rle_encode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.lo_rle_decode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    rng = np.random.default_rng(3)
    samples = [np.zeros(5, np.uint8), np.zeros(6, np.uint8), np.zeros(70000, np.uint8), rng.integers(0, 3, 5000).astype(np.uint8),
               np.repeat(rng.integers(0, 255, 400), rng.integers(1, 12, 400)).astype(np.uint8), np.array([9], np.uint8),
               np.concatenate([rng.integers(0, 255, 40000), np.zeros(40000)]).astype(np.uint8)]
    for s in samples:
        n = lib.lo_rle_size(s.ctypes.data, s.size)
        out = np.zeros(n + 8, np.uint8)
        assert lib.lo_rle_encode(s.ctypes.data, s.size, out.ctypes.data) == n
        back = np.full(s.size, 0x55, np.uint8)
        assert lib.lo_rle_decode(out.ctypes.data, n, back.ctypes.data, s.size) == 1
        assert np.array_equal(back, s)
    # a 5-run at the very end stays literal, a 6-run becomes a repeat token (RLE.cpp:74-79)
    five = np.zeros(5, np.uint8); six = np.zeros(6, np.uint8)
    assert lib.lo_rle_size(five.ctypes.data, 5) == 2 + 5 + 2 and lib.lo_rle_size(six.ctypes.data, 6) == 3 + 2


def test_oracle_rejects_and_accepts_corrupted_blobs_like_the_reference(oracle):
    """robustness pinning: 1-3 flipped bytes (checksum repaired so that the parsers are reached) -- the oracle's verdict and
    pixels equal the unmodified reference's.  The GPU fuzz test (tests/test_gpu_fuzz.py) then holds the product to the oracle."""
    from test_gpu_fuzz import repair
    from cases import c2_raster, c4_raster
    ref = ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libLerc_ref.so absent")
    rng = np.random.default_rng(5)
    for arr, mz, kw in [(c2_raster(40, 41), 0.01, {}), (c4_raster(48, 64), 0, {"n_depth": 3}),
                        (c2_raster(64, 64), 0.01, {"mask": (rng.random((64, 64)) > 0.2).astype(np.uint8)})]:
        st, blob, _ = oracle.encode(arr, mz, **kw)
        assert st == 0
        st, info = oracle.blob_info(blob)
        stricter = 0
        for _ in range(60):
            b = bytearray(blob)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(95, len(b)))] ^= int(rng.integers(1, 256))
            bad = repair(bytes(b))
            s_o, d_o, m_o = oracle.decode(bad, info=info)
            s_r, d_r, m_r = ref.decode(bad, info=info)
            if "mask" in kw and s_o != 0 and s_r == 0:
                # A corrupted MASK can give a block more valid pixels than its bit-stuffed value count.  The reference then reads
                # past the end of its value vector (Lerc2.cpp:2176-2199 has no bound check for version > 2: undefined behaviour,
                # stale values of earlier blocks); the oracle and the product reject the blob.  Documented deviation.
                stricter += 1
                continue
            assert (s_o == 0) == (s_r == 0)
            if s_r == 0:
                assert np.array_equal(d_o.view(np.uint8), d_r.view(np.uint8))
        assert stricter <= 6
