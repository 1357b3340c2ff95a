#!/bin/bash
# round 2 GPU call: default bench line with the c3 / c4 / c5 sub-records at N GPUs
set -u
N=${NGPU:-1}
OUT=gpurun_out/${R2OUT:-r2m}
mkdir -p "$OUT"
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
fi
tail -3 "$OUT/bench_n$N.err"
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_n$N.json") if l.startswith("{")][-1])
print("c2", round(d["value"], 2), "Gpx/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"], 2), {n: round(v["ms_per_step"], 4) for n, v in d["roofline"]["kernels"].items()})
for k in ("c5", "c4", "c3"):
    r = d.get(k)
    if r: print(k, json.dumps({x: r[x] for x in r if x in ("value", "ms_per_step", "error", "gather", "kernels", "step_frac", "per_gpu_gpixels")})[:900])
print("cpu", d.get("cpu_baseline"))
PY
