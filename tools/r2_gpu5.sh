#!/bin/bash
set -u
OUT=gpurun_out/r2n; mkdir -p $OUT
for v in default nofill; do
  if [ $v = nofill ]; then export LERC_B200_DBG_NOFILL=1; fi
  timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-sub > $OUT/b_$v.json 2> $OUT/b_$v.err
  python - <<PY
import json
d = json.load(open("$OUT/b_$v.json"))
print("$v", round(d["value"], 2), "Gpx/s", round(d["ms_per_step"], 4), "ms", {n: round(v["ms_per_step"], 4) for n, v in d["roofline"]["kernels"].items()})
PY
done
