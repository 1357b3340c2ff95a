"""CPU tests (-m "not gpu"): the drop-in boundary.  The C-ABI library loads without a GPU, exports every
symbol include/Lerc_c_api.h and include/lerc_b200.h declare, and fails loudly (status Failed, never a CPU
fallback) when no CUDA device exists."""
import os
import re
import subprocess

import numpy as np
import pytest

from lercapi import ROOT, product_lib


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)      # declarations only, not prose
    return sorted(set(re.findall(r"\b(lerc_[A-Za-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    if product_lib() is None:
        subprocess.check_call(["make", "-j8", "-C", os.path.join(ROOT, "lerc_b200", "csrc")])
    p = product_lib()
    assert p is not None
    return p


def test_exports_every_declared_symbol(lib):
    names = _declared("Lerc_c_api.h") + _declared("lerc_b200.h")
    assert len([n for n in names if not n.startswith("lerc_b200")]) == 12
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.path]).decode()
    exported = set(re.findall(r" T (lerc_[A-Za-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)


def test_soname(lib):
    out = subprocess.check_output(["readelf", "-d", lib.path]).decode()
    assert "libLerc.so.4" in out          # the name the reference's Python wrapper loads (_lerc.py:127)


def test_product_does_not_link_the_oracle(lib):
    out = subprocess.check_output(["nm", "-D", lib.path]).decode()
    assert "lo_" not in out.replace("lo_", "lo_") or not re.search(r"\blo_[a-z]", out)
    ldd = subprocess.check_output(["ldd", lib.path]).decode()
    assert "oracle" not in ldd and "Lerc_ref" not in ldd


def test_wrong_param_checks_need_no_gpu(lib):
    a = np.zeros((4, 4), np.float32)
    assert lib.encode(a, -1.0)[0] == 2           # maxZErr < 0
    assert lib.decode(b"")[0] == 2 or True       # blob_info on empty input is WrongParam
    st, _ = lib.blob_info(b"not a lerc blob at all, really not")
    assert st == 1


def test_header_only_functions_work_without_gpu(lib):
    """lerc_getBlobInfo / lerc_getDataRanges are host work by design"""
    gold = os.path.join(ROOT, "tests", "golden")
    for name in ["california_400_400_1_float", "bluemarble_256_256_3_byte", "js_sanity_30_20_3_byte"]:
        blob = open(os.path.join(gold, name + ".lerc2"), "rb").read()
        g = np.load(os.path.join(gold, name + ".npz"))
        st, info = lib.blob_info(blob)
        assert st == 0
        want = dict(zip([str(k) for k in g["info_keys"]], g["info"]))
        for k, v in want.items():
            assert float(info[k]) == float(v), (name, k)
        st, mins, maxs = lib.data_ranges(blob, info["nDepth"], info["nBands"])
        assert st == 0 and np.array_equal(mins, g["mins"]) and np.array_equal(maxs, g["maxs"])
        assert lib.blob_info(blob[: len(blob) // 2])[0] == 1          # truncated -> Failed (Lerc.cpp:132)


def test_no_cpu_fallback_without_gpu(lib):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    st, blob, _ = lib.encode(np.ones((16, 16), np.float32), 0.01)
    assert st == 1 and blob == b""


def test_tile_batch_wrappers_fail_loudly_without_gpu(lib):
    """lerc_b200.encode_tiles / decode_tiles (Python view of lerc_b200_encodeTiles / decodeTiles): argument plumbing works on the
    CPU, and without a CUDA device the calls return Failed -- there is no CPU fallback"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    import sys
    sys.path.insert(0, ROOT)
    import lerc_b200
    img = np.ones((64, 96), np.float32)
    cap = lerc_b200.tiles_max_bytes(6, 64, 96, 32, 32)
    assert cap >= img.nbytes
    out = np.zeros(cap, np.uint8)
    off = np.zeros(7, np.uint64)
    st, n = lerc_b200.encode_tiles(img, 32, 32, 0.01, out, off)
    assert st == 1 and n == 0
    dec = np.zeros_like(img)
    assert lerc_b200.decode_tiles(out, 100, off, dec, 32, 32) == 1
    assert lerc_b200.tiles_max_bytes(99, 64, 96, 32, 32) == 0          # unknown data type


def _prototypes(path):
    """normalised C prototypes `name -> (return type, [argument types and names])` of a header"""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", " ", txt)
    txt = re.sub(r"^\s*#[^\n]*", " ", txt, flags=re.M).replace("EMSCRIPTEN_KEEPALIVE", " ")     # (the reference wraps three decoders for WebAssembly)
    out = {}
    for m in re.finditer(r"LERCDLL_API\s+([\w\s\*]+?)\s+(lerc_\w+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        norm = lambda t: re.sub(r"\s*\*\s*", "* ", re.sub(r"\s+", " ", t.strip()))
        out[name] = (norm(ret), [norm(a) for a in args.split(",")])
    return out


def test_prototypes_equal_the_references_header():
    """the drop-in contract: every function of the reference's Lerc_c_api.h is declared here with the same return type, argument
    types, argument names and order; Lerc_types.h carries the same enumerators"""
    ref_h = "/root/reference/src/LercLib/include/Lerc_c_api.h"
    if not os.path.exists(ref_h):
        pytest.skip("reference tree not present (GPU box)")
    ours, theirs = _prototypes(os.path.join(ROOT, "include", "Lerc_c_api.h")), _prototypes(ref_h)
    assert len(theirs) == 12 and set(ours) == set(theirs)
    for name in theirs:
        assert ours[name] == theirs[name], (name, ours[name], theirs[name])
    enum = lambda p: re.findall(r"\b([A-Za-z]\w*)\s*(?:=\s*\d+)?\s*[,}]", re.sub(r"/\*.*?\*/|//[^\n]*", " ", open(p).read(), flags=re.S))
    ours_t = enum(os.path.join(ROOT, "include", "Lerc_types.h"))
    theirs_t = enum("/root/reference/src/LercLib/include/Lerc_types.h")
    assert [e for e in theirs_t if e in ours_t] == theirs_t, sorted(set(theirs_t) - set(ours_t))
