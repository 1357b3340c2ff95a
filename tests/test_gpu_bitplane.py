"""GPU tests (-m gpu): maxZErr == 777, the reference's "cheat code" for the integer bit-plane mode (Lerc2.cpp:210-217;
Lerc2::TryBitPlaneCompression, Lerc2.cpp:1071-1229): k_bitplane_counts + the reference's decision on the host raise maxZError to
half of the last bit plane that is not noise.  The oracle is pinned to the reference by test_oracle_vs_reference.py::test_bitplane_mode_hashes."""
import numpy as np
import pytest

from cases import bitplane_cases
from lercapi import encode_tiles, oracle_lib, product_lib, tile_windows

pytestmark = pytest.mark.gpu
CASES = bitplane_cases()


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib()
    assert prod is not None and orc is not None
    return prod, orc


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_bitplane_mode_matches_oracle(libs, case):
    prod, orc = libs
    name, arr, kw = case
    s_o, b_o, _ = orc.encode(arr, 777, **kw)
    s_p, b_p, _ = prod.encode(arr, 777, **kw)
    assert s_p == s_o, f"status {s_p} vs oracle {s_o}"
    if s_o != 0:
        return
    assert b_p == b_o, f"blob differs ({len(b_p)} vs {len(b_o)} bytes)"
    st, n = prod.compute_size(arr, 777, **kw)
    assert st == 0 and n == len(b_o)
    assert prod.blob_info(b_p)[1]["maxZErrUsed"] == orc.blob_info(b_o)[1]["maxZErrUsed"]
    t_o, d_o, _ = orc.decode(b_o)
    t_p, d_p, _ = prod.decode(b_o)
    assert t_o == 0 and t_p == 0 and np.array_equal(d_p.view(np.uint8), d_o.view(np.uint8))


def test_bitplane_mode_through_the_tile_batch(libs):
    prod, orc = libs
    arr = CASES[0][1]                                                      # 120 x 160 uint16 with six noise planes
    st, blobs, _ = encode_tiles(prod, arr, 80, 80, 777)
    assert st == 0
    for t, (ys, xs) in enumerate(tile_windows(arr.shape[0], arr.shape[1], 80, 80)):
        s_o, b_o, _ = orc.encode(np.ascontiguousarray(arr[ys, xs]), 777)
        assert s_o == 0 and blobs[t] == b_o, f"tile {t}"


def test_bitplane_mode_edge_cases(libs):
    """empty bands, float types and old codec versions with the cheat code (Lerc2.cpp:205-224; v6 floats: the noData / NaN filter
    resets maxZError for an empty band first, Lerc.cpp:1473-1478)"""
    prod, orc = libs
    rng = np.random.default_rng(0)
    a16 = rng.integers(0, 1000, (90, 100)).astype(np.int16)
    f32 = rng.random((90, 100)).astype(np.float32)
    zero = np.zeros((90, 100), np.uint8)
    for arr, kw in [(a16, dict(mask=zero)), (f32, dict(mask=zero)), (f32, dict(mask=zero, version=5)), (np.full((90, 100), np.nan, np.float32), {}),
                    (a16, dict(version=4)), (f32, {}), (a16[:20, :30].copy(), dict(version=3))]:
        s_o, b_o, _ = orc.encode(arr, 777, **kw)
        s_p, b_p, _ = prod.encode(arr, 777, **kw)
        assert s_p == s_o and b_p == b_o, kw
