#!/bin/bash
# round 2 GPU call: the 8-bit Huffman path (BASELINE config 4): tests, timing, one ncu capture per kernel of the step
set -u
OUT=gpurun_out/${R2OUT:-r2c4}
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --workload c4 --steps 5 --no-cpu-baseline > "$OUT/bench_c4.json" 2> "$OUT/bench_c4.err"
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_c4.json") if l.startswith("{")][-1])
r = d.get("record", d)
print("c4", round(d["value"], 2), "Gpx/s", r.get("ms_per_step"), json.dumps(r.get("kernels"))[:900])
PY
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_huffman_segments|k_huff_emit|k_huff_rows|k_huff_chunks|k_tiles|k_histograms" -c 8 -o "$OUT/prof_c4" python bench.py --workload c4 --steps 1 --no-cpu-baseline > "$OUT/ncu_c4.log" 2>&1
tail -2 "$OUT/ncu_c4.log"
fi
