#!/usr/bin/env python3
"""tools/cusim/check.py -- DEVELOPMENT TOOL, NOT PRODUCT, NOT A TEST OF THE PRODUCT.

Drives tools/cusim/_build/libLerc_sim.so (the product sources compiled against the host-side CUDA-semantics shim)
through the same C-ABI calls as tests/test_gpu_parity.py and compares with the oracle.  A pass here means the kernel
LOGIC is right for these inputs; the GPU parity tests (-m gpu, on the B200) remain the only parity evidence.

  python tools/cusim/check.py [-k substring] [--size H W] [--versions]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lercapi import LercLib, oracle_lib  # noqa: E402
from cases import all_cases  # noqa: E402


def sim_lib():
    return LercLib(os.path.join(ROOT, "tools", "cusim", os.environ.get("CUSIM_BUILD_DIR", "_build"), "libLerc_sim.so"))


def check_case(sim, orc, name, arr, mz, kw, version=None):
    t0 = time.time()
    kw = dict(kw, version=version)
    st_o, blob_o, _ = orc.encode(arr, mz, **kw)
    st_s, blob_s, buf = sim.encode(arr, mz, **kw)
    msg = []
    if st_o != st_s:
        msg.append(f"encode status {st_s} != oracle {st_o}")
    elif st_o == 0:
        if blob_s != blob_o:
            n = min(len(blob_s), len(blob_o))
            diff = next((i for i in range(n) if blob_s[i] != blob_o[i]), n)
            msg.append(f"blob differs (len {len(blob_s)} vs {len(blob_o)}, first diff at {diff})")
        if buf[len(blob_s):].any():
            msg.append("output buffer not zero-filled after the blob")
        st_o2, dec_o, mask_o = orc.decode(blob_o)
        st_s2, dec_s, mask_s = sim.decode(blob_o)
        if st_o2 != st_s2:
            msg.append(f"decode status {st_s2} != oracle {st_o2}")
        elif st_o2 == 0:
            if not np.array_equal(dec_s.view(np.uint8), dec_o.view(np.uint8)):
                msg.append("decoded pixels differ")
            if (mask_o is None) != (mask_s is None) or (mask_o is not None and not np.array_equal(mask_o, mask_s)):
                msg.append("decoded mask differs")
        st_c, n_c = sim.compute_size(arr, mz, **kw)
        if st_c != 0 or n_c != len(blob_o):
            msg.append(f"computeCompressedSize {n_c} (status {st_c}) != {len(blob_o)}")
    print(f"{'ok  ' if not msg else 'FAIL'} {name:34s} {time.time() - t0:6.1f}s  {'; '.join(msg)}", flush=True)
    return not msg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-k", default="")
    ap.add_argument("--size", nargs=2, type=int, default=[70, 90])
    ap.add_argument("--versions", action="store_true", help="also lerc_encodeForVersion with codec versions 2..5")
    a = ap.parse_args()
    sim, orc = sim_lib(), oracle_lib()
    bad = 0
    for version in ([None, 2, 3, 4, 5] if a.versions else [None]):
        for name, arr, mz, kw in all_cases(a.size[0], a.size[1]):
            if a.k in name:
                bad += not check_case(sim, orc, name + ("" if version is None else f" v{version}"), arr, mz, kw, version)
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
