#!/bin/bash
set -u
OUT=gpurun_out/r2p; mkdir -p $OUT
for v in default noraise nofinish noraise,nofinish; do
  LERC_B200_DBG=$v timeout 200 python bench.py --steps 10 --no-cpu-baseline --no-sub > $OUT/b_$v.json 2> $OUT/b_$v.err
  python - <<PY
import json
try:
  d = json.load(open("$OUT/b_$v.json"))
  print("$v", round(d["value"], 2), "Gpx/s", round(d["ms_per_step"], 4), "ms", {n: round(v["ms_per_step"], 4) for n, v in d["roofline"]["kernels"].items() if "encode_tile" in n})
except Exception as e: print("$v", "ERR", e, open("$OUT/b_$v.err").read()[-300:])
PY
done
