#!/usr/bin/env python3
"""tools/cusim/check_tiles.py -- DEVELOPMENT TOOL (see check.py): the tile batch entry points on the simulator build.
Every tile's blob must equal the oracle's lerc_encode of that window; the batch decode must equal the oracle's decode."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lercapi import LercLib, oracle_lib, encode_tiles, decode_tiles, tile_windows  # noqa: E402
from cases import tile_cases  # noqa: E402


def main():
    sim = LercLib(os.path.join(ROOT, "tools", "cusim", os.environ.get("CUSIM_BUILD_DIR", "_build"), "libLerc_sim.so"))
    orc = oracle_lib()
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    bad = 0
    for name, raster, tr, tc, mz in tile_cases():
        if only not in name:
            continue
        t0 = time.time()
        msg = []
        st, blobs, off = encode_tiles(sim, raster, tr, tc, mz)
        wins = list(tile_windows(raster.shape[0], raster.shape[1], tr, tc))
        if st:
            msg.append(f"encodeTiles status {st}")
        else:
            want = []
            for t, (ys, xs) in enumerate(wins):
                st_o, blob_o, _ = orc.encode(np.ascontiguousarray(raster[ys, xs]), mz)
                assert st_o == 0
                want.append(blob_o)
                if blobs[t] != blob_o and len(msg) < 4:
                    n = min(len(blobs[t]), len(blob_o))
                    diff = next((i for i in range(n) if blobs[t][i] != blob_o[i]), n)
                    msg.append(f"tile {t}: blob differs (len {len(blobs[t])} vs {len(blob_o)}, first diff {diff})")
            st, dec = decode_tiles(sim, want, raster.dtype, raster.shape[0], raster.shape[1], tr, tc)
            if st:
                msg.append(f"decodeTiles status {st}")
            else:
                for t, (ys, xs) in enumerate(wins):
                    st_o, dec_o, _ = orc.decode(want[t])
                    if not np.array_equal(dec[ys, xs].view(np.uint8), dec_o[0, :, :, 0].view(np.uint8)) and len(msg) < 8:
                        msg.append(f"tile {t}: decoded pixels differ")
        print(f"{'ok  ' if not msg else 'FAIL'} {name:34s} {len(wins):4d} tiles {time.time() - t0:6.1f}s  {'; '.join(msg)}", flush=True)
        bad += bool(msg)
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
