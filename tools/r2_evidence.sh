#!/bin/bash
# round 2 evidence call (one GPU): whole GPU suite, the default bench line, the reference arm, the ncu launch list of the same bench
# command, and one --set full capture per dominant kernel (DRAM traffic per launch -> profiles/dram_traffic.json by tools/r2_collect.py)
set -u
OUT=gpurun_out/${R2OUT:-r2ev}
mkdir -p "$OUT"
git rev-parse HEAD > "$OUT/commit.txt" 2>/dev/null || cp .commit_for_gpu "$OUT/commit.txt" 2>/dev/null
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > "$OUT/gpu_tests.log" 2>&1
tail -3 "$OUT/gpu_tests.log"
timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference_arm.json" 2> "$OUT/bench_reference_arm.err"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1
LERC_B200_STRIP_LOG2=30 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_encode_tile|k_decode_stream" -c 2 -o "$OUT/prof_c2" python tools/big_check.py > "$OUT/ncu_c2.log" 2>&1
tail -2 "$OUT/ncu_c2.log"
# (the reports are too large to travel: export what profiles/ keeps, then drop them)
export_rep() { ncu -i "$OUT/$1.ncu-rep" --page raw --csv > "$OUT/$1_raw.csv" 2>/dev/null; ncu -i "$OUT/$1.ncu-rep" --page details > "$OUT/$1_details.txt" 2>/dev/null; }
export_rep prof_c2
for k in k_encode_tile k_decode_stream; do ncu -i "$OUT/prof_c2.ncu-rep" --page source --csv --kernel-name regex:$k > "$OUT/prof_c2_source_$k.csv" 2>/dev/null; done
rm -f "$OUT/prof_c2.ncu-rep"
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_n1.json") if l.startswith("{")][-1])
print("c2", round(d["value"], 2), "Gpx/s", round(d["ms_per_step"], 4), "ms step_frac", round(d["roofline"]["step_frac"], 3), "e2e", round(d["e2e"]["value"], 2), d["e2e"].get("two_callers", {}).get("value"), d["e2e"].get("pcie_measured_gbs"))
for k in ("c5", "c4", "c3"):
    r = d.get(k)
    if r: print(k, json.dumps({x: r[x] for x in r if x in ("value", "ms_per_step", "error", "step_frac")}))
print("cpu", d.get("cpu_baseline"))
print(open("$OUT/bench_reference_arm.json").read()[:600])
PY
# tile batch (config 5 shape) and 8-bit Huffman (config 4) kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_encode_tile|k_tiles_blocks|k_tiles_finish|k_tiles_parse" -c 4 -o "$OUT/prof_c5" python bench.py --workload c5 --steps 1 --no-cpu-baseline > "$OUT/ncu_c5.log" 2>&1
tail -1 "$OUT/ncu_c5.log"
export_rep prof_c5; rm -f "$OUT/prof_c5.ncu-rep"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_huffman_runs|k_huff_emit|k_huff_rows|k_huff_chunks|k_tiles_count8|k_histograms|k_stats" -c 10 -o "$OUT/prof_c4" python bench.py --workload c4 --steps 1 --no-cpu-baseline > "$OUT/ncu_c4.log" 2>&1
tail -1 "$OUT/ncu_c4.log"
export_rep prof_c4; rm -f "$OUT/prof_c4.ncu-rep"
du -sh "$OUT"
for w in c4 c5 c3 c2m; do timeout 600 python bench.py --workload $w --steps 5 --no-cpu-baseline > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"; done
