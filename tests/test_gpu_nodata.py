"""GPU tests (-m gpu): the _4D entry points with per-band noData values (Lerc_c_api.h:295-380).
Encode: Lerc::FilterNoDataAndNaN / FilterNoData (Lerc.cpp:1378-1618, :1241-1374) run as device kernels + host decisions
(lerc_encode.cu:prefilterNoData); decode: pUsesNoData / noDataValues outputs and Lerc::RemapNoData (Lerc.cpp:1046-1076).
The oracle is pinned to the reference on these cases by tests/test_oracle_vs_reference.py::test_nodata_4d_hashes."""
import numpy as np
import pytest

from cases import nodata_cases
from lercapi import oracle_lib, product_lib

pytestmark = pytest.mark.gpu
CASES = nodata_cases()


@pytest.fixture(scope="module")
def libs():
    prod, orc = product_lib(), oracle_lib()
    assert prod is not None and orc is not None
    return prod, orc


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_nodata_matches_oracle(libs, case):
    prod, orc = libs
    name, arr, mz, kw = case
    s_o, b_o = orc.encode_4d(arr, mz, **kw)
    s_p, b_p = prod.encode_4d(arr, mz, **kw)
    assert s_p == s_o, f"status {s_p} vs oracle {s_o}"
    if s_o != 0:
        return
    assert b_p == b_o, f"blob differs ({len(b_p)} vs {len(b_o)} bytes)"
    assert prod.encode_4d(arr, mz, size_only=True, **kw) == (0, len(b_o))
    o = orc.decode_4d(b_o)
    p = prod.decode_4d(b_o)
    assert o[0] == 0 and p[0] == 0
    assert np.array_equal(p[1].view(np.uint8), o[1].view(np.uint8)), "decoded pixels differ"
    assert (o[2] is None) == (p[2] is None) and (o[2] is None or np.array_equal(o[2], p[2]))
    assert np.array_equal(p[3], o[3]) and np.array_equal(p[4], o[4]), "pUsesNoData / noDataValues differ"
    assert prod.decode_4d(b_o, want_no_data=False)[0] == orc.decode_4d(b_o, want_no_data=False)[0]      # HasNoData when the caller would miss it
    od, pd = orc.decode_4d(b_o, to_double=True), prod.decode_4d(b_o, to_double=True)
    assert od[0] == 0 and pd[0] == 0 and np.array_equal(pd[1].view(np.uint8), od[1].view(np.uint8))
