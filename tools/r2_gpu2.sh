#!/bin/bash
# round 2 GPU call: parity of the tile encoder + stream decoder, timing, ncu captures of both
set -u
OUT=gpurun_out/${R2OUT:-r2f}
mkdir -p "$OUT"
timeout 400 python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_fuzz.py tests/test_gpu_tiles.py -x -q -m gpu -p no:cacheprovider > "$OUT/tests.log" 2>&1
tail -3 "$OUT/tests.log"
timeout 200 python bench.py --steps 20 --no-cpu-baseline > "$OUT/bench_c2.json" 2> "$OUT/bench_c2.err"
python - <<'PY'
import json, os
d = json.load(open("gpurun_out/" + os.environ.get("R2OUT", "r2f") + "/bench_c2.json"))
print(round(d["value"], 2), "Gpx/s", round(d["ms_per_step"], 4), "ms", {n: round(v["ms_per_step"], 4) for n, v in d["roofline"]["kernels"].items()})
print("e2e", d["e2e"]["value"], d["clocks"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_encode_tile|k_decode_stream" -c 2 -o "$OUT/prof" python tools/big_check.py > "$OUT/ncu.log" 2>&1
tail -5 "$OUT/ncu.log"
