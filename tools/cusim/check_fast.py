#!/usr/bin/env python3
"""tools/cusim/check_fast.py -- DEVELOPMENT TOOL, NOT PRODUCT, NOT A TEST OF THE PRODUCT.

The single-pass encoder (lerc_encode_tile.cuh) and the single-kernel decoder on the simulator build: shapes around the
tile geometry (several tiles per block row, partial tiles, rows that are not 16-byte multiples, values wider than 16 bits =
several packing passes per tile), all pixel types, against the oracle; asserts through the simulator library's own
counters that the fast paths were taken.

  python tools/cusim/check_fast.py [-k substring]
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lercapi import LercLib, oracle_lib  # noqa: E402
from cases import smooth_field  # noqa: E402


def sim_lib():
    return LercLib(os.path.join(ROOT, "tools", "cusim", os.environ.get("CUSIM_BUILD_DIR", "_build"), "libLerc_sim.so"))


def stats(sim):
    out = (ctypes.c_ulonglong * 5)()
    sim.lib.lerc_b200_get_stats(out, 5)
    return list(out)


def make(dtype, h, w, seed, amp=0.5):
    rng = np.random.default_rng(seed)
    base = smooth_field(h, w) + rng.normal(0, amp, (h, w))
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        return np.clip(base * 3 - (2000 if info.min < 0 else 0), info.min, info.max).astype(dtype)
    return base.astype(dtype)


def cases():
    shapes = [(8, 8), (64, 64), (257, 300), (5, 1000), (1000, 5), (64, 1024), (24, 1032), (16, 2048), (17, 2049), (9, 4104), (40, 3000)]
    kinds = [(np.float32, 0.01), (np.float32, 1.0), (np.float64, 0.001), (np.int16, 0), (np.uint16, 2), (np.int32, 0), (np.uint32, 3)]
    for h, w in shapes:
        for dt, mz in kinds:
            yield f"{np.dtype(dt).name}_{mz}_{h}x{w}", make(dt, h, w, h * 7919 + w), mz, True
    # values wider than 16 bits: generic blocks, more than STAGE_CAP bytes per tile -> several passes
    yield "f32_wide_1e-5_24x2056", make(np.float32, 24, 2056, 3, amp=30.0), 1e-5, True
    yield "f64_wide_1e-7_24x1100", make(np.float64, 24, 1100, 4, amp=30.0), 1e-7, True
    yield "i32_wide_lossless_16x3000", (make(np.int32, 16, 3000, 5).astype(np.int64) * 30011 % (1 << 27)).astype(np.int32), 0, True
    # flat areas (zero / constant blocks) next to noisy ones
    sea = make(np.float32, 64, 2200, 6); sea[:24, :] = 0; sea[30:50, 100:1900] = 37.0
    yield "f32_sea_lake_64x2200", sea, 0.01, None
    # mixed: integer-valued region (equal neighbours, LUT filter path) + noise
    mix = make(np.float32, 48, 1500, 7); mix[:, :700] = np.round(mix[:, :700] / 8) * 8
    yield "f32_mixed_steps_48x1500", mix, 0.01, None


def main():
    os.environ.setdefault("LERC_B200_VERBOSE", "1")
    ap = argparse.ArgumentParser()
    ap.add_argument("-k", default="")
    args = ap.parse_args()
    sim, orc = sim_lib(), oracle_lib()
    fails = 0
    for name, arr, mz, must_fast in cases():
        if args.k not in name:
            continue
        t0 = time.time()
        s_o, b_o, _ = orc.encode(arr, mz)
        before = stats(sim)
        s_s, b_s, _ = sim.encode(arr, mz)
        mid = stats(sim)
        msg = []
        if s_o != 0 or s_s != 0:
            msg.append(f"status {s_s} / oracle {s_o}")
        elif b_s != b_o:
            n = min(len(b_s), len(b_o))
            diff = next((i for i in range(n) if b_s[i] != b_o[i]), n)
            msg.append(f"blob differs (len {len(b_s)} vs {len(b_o)}, first diff at {diff})")
        if must_fast and mid[3] != before[3] + 1:
            msg.append("single-pass encoder not taken")
        if s_o == 0:
            t_o, d_o, _ = orc.decode(b_o)
            # stderr of the library (LERC_B200_VERBOSE names a decoder that gave up) goes through a temporary file
            import tempfile
            sys.stderr.flush()
            saved = os.dup(2)
            with tempfile.TemporaryFile() as tf:
                os.dup2(tf.fileno(), 2)
                try:
                    t_s, d_s, _ = sim.decode(b_o)
                finally:
                    os.dup2(saved, 2); os.close(saved)
                tf.seek(0)
                if os.environ.get("DS_DEBUG"):
                    sys.stdout.write(tf.read().decode()); tf.seek(0)
                note = " ".join(l.decode().strip().replace("[lerc_b200] ", "") for l in tf.readlines() if b"status" in l)
            after = stats(sim)
            if t_s != t_o or (t_o == 0 and not np.array_equal(d_s.view(np.uint8), d_o.view(np.uint8))):
                msg.append("decode differs")
            fast_dec = after[4] == mid[4] + 1
        else:
            fast_dec = False
        print(f"{'ok  ' if not msg else 'FAIL'} {name:36s} {time.time() - t0:6.1f}s enc_fast={mid[3] - before[3]} dec_fast={int(fast_dec)} {note if s_o == 0 else ''} {'; '.join(msg)}", flush=True)
        fails += bool(msg)
    print("failures:", fails)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
