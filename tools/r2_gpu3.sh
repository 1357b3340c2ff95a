#!/bin/bash
# round 2 GPU call: whole GPU suite + bench of the default workload
set -u
OUT=gpurun_out/${R2OUT:-r2k}
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > "$OUT/tests.log" 2>&1
tail -5 "$OUT/tests.log"
timeout 300 python bench.py --steps 20 --no-cpu-baseline > "$OUT/bench_c2.json" 2> "$OUT/bench_c2.err"
python - <<'PY'
import json, os
d = json.load(open("gpurun_out/" + os.environ.get("R2OUT", "r2k") + "/bench_c2.json"))
print(round(d["value"], 2), "Gpx/s", round(d["ms_per_step"], 4), "ms", {n: round(v["ms_per_step"], 4) for n, v in d["roofline"]["kernels"].items()})
print("e2e", d["e2e"]["value"], d["clocks"])
PY
