#!/usr/bin/env python3
"""tools/cusim/check_fpl.py -- DEVELOPMENT TOOL (see check.py): the lossless float (FPL) encoder on the simulator build.
Float rasters at maxZError 0: blob == the oracle's (which is pinned to the reference's blobs), size query == blob size, and the
simulator build decodes its own blob back to the input."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lercapi import LercLib, oracle_lib  # noqa: E402
from cases import fpl_cases, fpl_encode_cases  # noqa: E402


def main():
    sim = LercLib(os.path.join(ROOT, "tools", "cusim", os.environ.get("CUSIM_BUILD_DIR", "_build"), "libLerc_sim.so"))
    orc = oracle_lib(fpl_encoder=True)
    only = sys.argv[1:]
    bad = 0
    for name, arr, kw in fpl_cases() + fpl_encode_cases():
        if only and name not in only:
            continue
        t0 = time.time()
        s_o, b_o, _ = orc.encode(arr, 0.0, **kw)
        s_s, b_s, _ = sim.encode(arr, 0.0, **kw)
        msg = []
        if s_o != s_s:
            msg.append(f"status {s_s} vs oracle {s_o}")
        elif b_s != b_o:
            n = min(len(b_s), len(b_o))
            msg.append(f"blob differs (len {len(b_s)} vs {len(b_o)}, first diff {next((i for i in range(n) if b_s[i] != b_o[i]), n)})")
        elif sim.compute_size(arr, 0.0, **kw) != (0, len(b_o)):
            msg.append("size query differs")
        else:
            st, data, _ = sim.decode(b_s)
            t_o, d_o, _ = orc.decode(b_s)
            if st or t_o or not np.array_equal(data.view(np.uint8), d_o.view(np.uint8)):
                msg.append("decode differs")
        bad += bool(msg)
        print(f"{'FAIL' if msg else 'ok  '} {name:24s} {len(b_o) if b_o else 0:9d} B  {time.time() - t0:5.1f}s  {'; '.join(msg)}")
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
