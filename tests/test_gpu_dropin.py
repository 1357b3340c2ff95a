"""Drop-in proofs and boundary hardening on the GPU:

* the reference's OWN callers run against the product library: `src/LercTest/main.cpp` (compiled by `make -C oracle ref` into
  oracle/_ref/lerctest_product, linked against lerc_b200/libLerc.so.4) and the reference's Python wrapper
  (`OtherLanguages/Python/lerc/_lerc.py:test()`, staged in oracle/_ref/pylerc) with the product library dropped beside it;
* BASELINE configs[0]: testData/california decode -> re-encode at its own maxZError and at 0 -> decode; testData/world.lerc1
  ingested through the reference (oracle/_ref/world_lerc1.npz) and round-tripped losslessly through the product;
* `Lerc::CheckDimensions` / `Lerc2::ReadHeader` guards (Lerc.cpp:1622-1639, Lerc2.cpp:877-911): status 6 and header-field fuzz
  against the oracle;
* concurrent callers (the reference is stateless and re-entrant, Lerc_c_api.h:111-124): four host threads encode / decode
  different rasters at once while another thread keeps the GPU busy with foreign kernels.
"""
import ctypes as C
import os
import shutil
import struct
import subprocess
import sys
import threading

import numpy as np
import pytest

from cases import c2_raster, smooth_field
from lercapi import ROOT, fletcher32, oracle_lib, product_lib, ref_lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
REFDIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib()
    assert prod is not None and orc is not None
    return prod, orc


# ---------------------------------------------------------------------------------------------------------------------
def test_reference_lerctest_program_against_product():
    exe = os.path.join(REFDIR, "lerctest_product")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/lerctest_product absent (make -C oracle ref)")
    env = dict(os.environ, LERCTEST_NONINTERACTIVE="1")
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    assert "SUMMARY: all good." in out, out[-2000:]
    # the binary really ran against the product (its dynamic section names libLerc.so.4 and the rpath points into lerc_b200/)
    dyn = subprocess.run(["readelf", "-d", exe], capture_output=True, text=True).stdout
    assert "libLerc.so.4" in dyn and "lerc_b200" in dyn


def test_reference_python_wrapper_against_product(tmp_path):
    src = os.path.join(REFDIR, "pylerc", "lerc")
    if not os.path.exists(os.path.join(src, "_lerc.py")):
        pytest.skip("oracle/_ref/pylerc absent (make -C oracle ref)")
    pkg = tmp_path / "lerc"
    shutil.copytree(src, pkg)
    shutil.copy(os.path.join(ROOT, "lerc_b200", "libLerc.so.4"), pkg / "libLerc.so.4")      # _lerc.py:127 loads exactly this name from its own directory
    code = "import sys, lerc; r = lerc.test(); print('lerc.test() ->', r); sys.exit(0 if r == 0 else 1)"
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, env=dict(os.environ, PYTHONPATH=str(tmp_path)), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]


# ---------------------------------------------------------------------------------------------------------------------
def test_config1_california_decode_reencode_decode(libs):
    """BASELINE configs[0] on the GPU: the shipped blob decodes to the reference's pixels; re-encoded at the blob's own
    maxZError (7.5e-05) and at 0 the product writes the reference's bytes; those decode back within the bound / exactly."""
    prod, orc = libs
    blob = open(os.path.join(GOLD, "california_400_400_1_float.lerc2"), "rb").read()
    gold = np.load(os.path.join(GOLD, "california_400_400_1_float.npz"))
    st, px, mask = prod.decode(blob)
    assert st == 0
    s_o, px_o, mask_o = orc.decode(blob)
    assert s_o == 0 and np.array_equal(px.view(np.uint8), px_o.view(np.uint8)) and np.array_equal(mask, mask_o)
    key = [k for k in gold.files if "pix" in k or "data" in k]
    if key:
        assert np.array_equal(np.asarray(gold[key[0]]).reshape(px.shape).view(np.uint8), px.view(np.uint8))
    st, info = prod.blob_info(blob)
    mz_blob = info.get("maxZErrUsed", 7.5e-05)
    img, m = px[0, :, :, 0], mask[0] if mask.ndim == 3 else mask
    ref = ref_lib()
    for mz in (mz_blob, 0.0):
        s_p, b_p, _ = prod.encode(img, mz, mask=m)
        s_o, b_o, _ = orc.encode(img, mz, mask=m)
        assert s_p == 0 and s_o == 0 and b_p == b_o, f"maxZError {mz}: {len(b_p)} vs {len(b_o)} bytes"
        if ref is not None and mz > 0:                             # (at 0 the reference leaves 4 uninitialised bytes per FPL plane: fpl_normalize)
            s_r, b_r, _ = ref.encode(img, mz, mask=m)
            assert s_r == 0 and b_r == b_p
        s2, px2, m2 = prod.decode(b_p)
        assert s2 == 0 and np.array_equal(m2.reshape(m.shape), m)
        valid = m.astype(bool)
        err = np.abs(px2[0, :, :, 0].astype(np.float64) - img.astype(np.float64))[valid].max()
        assert err <= mz * 1.1 + 1e-12 if mz > 0 else err == 0


def test_world_lerc1_through_reference_then_product(libs):
    """testData/world.lerc1 is a Lerc1 blob (out of scope for the product, which returns Failed for it): the reference decodes it
    (oracle/_ref/world_lerc1.npz, `make -C oracle ref`), the product codes those pixels + mask losslessly and at the file's scale."""
    prod, orc = libs
    p = os.path.join(REFDIR, "world_lerc1.npz")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/world_lerc1.npz absent (make -C oracle ref)")
    z = np.load(p)
    px, mask = z["pixels"], z["mask"]
    img = np.ascontiguousarray(px[0, :, :, 0])
    m = np.ascontiguousarray(mask.reshape(img.shape)) if mask.size else None
    for mz in (0.0, 0.5):
        s_p, b_p, _ = prod.encode(img, mz, mask=m)
        s_o, b_o, _ = orc.encode(img, mz, mask=m)
        assert s_p == 0 and s_o == 0 and b_p == b_o
        s2, px2, m2 = prod.decode(b_p)
        assert s2 == 0
        valid = m.astype(bool) if m is not None else np.ones(img.shape, bool)
        if m is not None:
            assert np.array_equal(m2.reshape(m.shape), m)
        d = np.abs(px2[0, :, :, 0].astype(np.float64) - img.astype(np.float64))[valid]
        assert (d == 0).all() if mz == 0 else d.max() <= mz * 1.1


# ---------------------------------------------------------------------------------------------------------------------
def test_dimensions_too_large_is_status_6(libs):
    """Lerc::CheckDimensions (Lerc.cpp:1622-1639) before any pixel is touched: same status as the oracle for every entry point"""
    prod, orc = libs
    a = np.zeros(64, np.float32)
    n = C.c_uint(0)
    out = np.zeros(256, np.uint8)
    for cols, rows, depth, dt in ((65536, 65536, 1, 6), (46341, 46341, 1, 1), (40000, 40000, 2, 6), (20000, 20000, 1, 7), (30000, 30000, 3, 1)):
        for lib in (prod, orc):
            st_s = lib.f["computeCompressedSize"](a.ctypes.data, dt, depth, cols, rows, 1, 0, None, 0.01, C.addressof(n))
            st_e = lib.f["encode"](a.ctypes.data, dt, depth, cols, rows, 1, 0, None, 0.01, out.ctypes.data, out.size, C.addressof(n))
            assert st_s == 6 and st_e == 6, (lib.path, cols, rows, depth, dt, st_s, st_e)
    # decode: the caller's dimensions are checked after the header matches, so a huge request against a small blob is Failed (1) on both
    st, blob, _ = orc.encode(c2_raster(16, 16), 0.01)
    buf = np.frombuffer(blob, np.uint8)
    for lib in (prod, orc):
        st = lib.f["decode"](buf.ctypes.data, len(blob), 0, None, 1, 65536, 65536, 1, 6, a.ctypes.data)
        assert st in (1, 6)
    s_p = prod.f["decode"](buf.ctypes.data, len(blob), 0, None, 1, 65536, 65536, 1, 6, a.ctypes.data)
    s_o = orc.f["decode"](buf.ctypes.data, len(blob), 0, None, 1, 65536, 65536, 1, 6, a.ctypes.data)
    assert s_p == s_o


def _repair(b):
    b = bytearray(b)
    size = struct.unpack_from("<i", b, 34)[0]
    if 14 < size <= len(b):
        struct.pack_into("<I", b, 10, fletcher32(bytes(b[14:size])))
    return bytes(b)


def test_header_field_fuzz_agrees_with_oracle(libs):
    """Lerc2::ReadHeader guards (Lerc2.cpp:877-911): every header field of a v6 blob set to hostile values (checksum repaired so
    that the guards, not the checksum, decide): getBlobInfo / decode statuses and, where both accept, pixels equal the oracle's."""
    prod, orc = libs
    img = c2_raster(40, 56)
    st, blob, _ = orc.encode(img, 0.01)
    assert st == 0
    fields = {"version": 6, "nRows": 14, "nCols": 18, "nDepth": 22, "numValidPixel": 26, "microBlockSize": 30, "blobSize": 34, "dataType": 38, "nBlobsMore": 42}
    hostile = [0, -1, 1, 2, 7, 8, 9, 33, 64, 255, 40, 56, 41, 2240, 2241, 2 ** 31 - 1, -(2 ** 31), len(blob) - 1, len(blob) + 1]
    st0, info0 = orc.blob_info(blob)
    n_cases = n_both_ok = 0
    for name, at in fields.items():
        for v in hostile:
            b = bytearray(blob)
            struct.pack_into("<i", b, at, v)
            bad = _repair(bytes(b)) if name != "version" else bytes(b)
            s_io, i_o = orc.blob_info(bad)
            s_ip, i_p = prod.blob_info(bad)
            assert s_ip == s_io, f"getBlobInfo {name}={v}: {s_ip} vs oracle {s_io}"
            if s_io == 0:
                assert i_p == i_o, f"getBlobInfo {name}={v}: info differs"
            s_o, d_o, m_o = orc.decode(bad, info=info0)
            s_p, d_p, m_p = prod.decode(bad, info=info0)
            assert (s_p == 0) == (s_o == 0), f"decode {name}={v}: status {s_p} vs oracle {s_o}"
            if s_o == 0:
                n_both_ok += 1
                assert np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8)), f"decode {name}={v}: pixels differ"
            n_cases += 1
    # doubles of the header: maxZError / zMin / zMax (NaN, Inf, negative, swapped)
    for at in (50, 58, 66):
        for v in (float("nan"), float("inf"), -float("inf"), -1.0, 0.0, 1e300, -1e300):
            b = bytearray(blob)
            struct.pack_into("<d", b, at, v)
            bad = _repair(bytes(b))
            s_o, d_o, _ = orc.decode(bad, info=info0)
            s_p, d_p, _ = prod.decode(bad, info=info0)
            assert (s_p == 0) == (s_o == 0), f"decode double@{at}={v}: status {s_p} vs oracle {s_o}"
            if s_o == 0:
                assert np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8)), f"decode double@{at}={v}: pixels differ"
            n_cases += 1
    assert n_cases > 150 and n_both_ok > 5


# ---------------------------------------------------------------------------------------------------------------------
def test_concurrent_callers(libs):
    """four host threads through the C ABI at once, different rasters / types / masks, while a fifth thread keeps the GPU busy with
    foreign kernels on its own stream: every blob equals the oracle's, every decode the oracle's pixels"""
    import torch
    prod, orc = libs
    rng = np.random.default_rng(11)
    jobs = []
    for k in range(4):
        h, w = 512 + 64 * k, 768 + 40 * k
        base = smooth_field(h, w) + rng.normal(0, 0.5, (h, w))
        if k == 0:
            jobs.append((base.astype(np.float32), 0.01, None))
        elif k == 1:
            jobs.append((np.clip(base * 3, -32768, 32767).astype(np.int16), 0, None))
        elif k == 2:
            m = np.ones((h, w), np.uint8); m[10:200, 30:400] = 0
            jobs.append((base.astype(np.float32), 0.1, m))
        else:
            jobs.append((np.clip(base / 6, 0, 255).astype(np.uint8), 0, None))
    want = []
    for arr, mz, m in jobs:
        s, b, _ = orc.encode(arr, mz, mask=m)
        assert s == 0
        s, d, _ = orc.decode(b)
        want.append((b, d))
    stop = threading.Event()
    errors = []

    def foreign():
        x = torch.randn(2048, 2048, device="cuda")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            while not stop.is_set():
                x = (x @ x).tanh()
        torch.cuda.synchronize()

    def worker(k):
        try:
            arr, mz, m = jobs[k]
            for _ in range(12):
                s, b, _ = prod.encode(arr, mz, mask=m)
                if s != 0 or b != want[k][0]:
                    errors.append(f"thread {k}: encode status {s}, blob equal {b == want[k][0]}")
                    return
                s, d, _ = prod.decode(b)
                if s != 0 or not np.array_equal(d.view(np.uint8), want[k][1].view(np.uint8)):
                    errors.append(f"thread {k}: decode status {s} or pixels differ")
                    return
        except Exception as ex:   # noqa: BLE001
            errors.append(f"thread {k}: {type(ex).__name__}: {ex}")

    f = threading.Thread(target=foreign)
    f.start()
    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    stop.set()
    f.join()
    assert not errors, errors


def test_concurrent_pipelined_host_calls(libs):
    """three host threads with HOST rasters large enough for the strip pipeline (copy-in / kernels / copy-out streams per pooled context;
    lerc_encode.cu encodeBandFast, lerc_decode.cu decodeStreamFast), pageable memory, different sizes: blobs and pixels equal the oracle's"""
    prod, orc = libs
    rng = np.random.default_rng(23)
    jobs, want = [], []
    for k in range(3):
        h, w = 1536 + 256 * k, 2048 + 8 * k                      # 12-17 MB: 8-11 strips
        arr = (smooth_field(h, w) + rng.normal(0, 0.5, (h, w))).astype(np.float32)
        s, b, _ = orc.encode(arr, 0.01)
        assert s == 0
        s, d, _ = orc.decode(b)
        jobs.append(arr); want.append((b, d))
    errors = []

    def worker(k):
        try:
            for _ in range(6):
                s, b, _ = prod.encode(jobs[k], 0.01)
                if s != 0 or b != want[k][0]:
                    errors.append(f"thread {k}: encode status {s}, blob equal {b == want[k][0]}")
                    return
                s, d, _ = prod.decode(b)
                if s != 0 or not np.array_equal(d.view(np.uint8), want[k][1].view(np.uint8)):
                    errors.append(f"thread {k}: decode status {s} or pixels differ")
                    return
        except Exception as ex:   # noqa: BLE001
            errors.append(f"thread {k}: {type(ex).__name__}: {ex}")

    th = [threading.Thread(target=worker, args=(k,)) for k in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
