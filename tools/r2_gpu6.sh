#!/bin/bash
# round 2 GPU call: e2e strip pipeline -- quick tests, then e2e for several strip sizes
set -u
OUT=gpurun_out/${R2OUT:-r2t}
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_large.py -q -m gpu -x -p no:cacheprovider > "$OUT/tests.log" 2>&1
tail -3 "$OUT/tests.log"
for L in 30 24 23 22 21; do
  LERC_B200_STRIP_LOG2=$L timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-sub > "$OUT/bench_c2_$L.json" 2> "$OUT/bench_c2_$L.err"
  python - <<PY
import json
d = json.load(open("$OUT/bench_c2_$L.json"))
print($L, round(d["value"], 2), "Gpx/s", "e2e", round(d["e2e"]["value"],3), d["e2e"])
PY
done
