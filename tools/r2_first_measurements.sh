#!/bin/bash
# tools/r2_first_measurements.sh -- first GPU call of round 2: parity + timing of the experimental kernel variants written (and checked
# on tools/cusim) at the end of round 1, against the defaults.  Run as:
#   gpurun --timeout 600 -- 'bash tools/r2_first_measurements.sh'
# Results land in gpurun_out/r2_first/ (one bench JSON line per configuration + the pytest tails).
set -u
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
run_bench() {   # name, env assignments..., then bench arguments after --
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 120 python bench.py "$@" --no-cpu-baseline > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
}
# 1. parity of the whole GPU suite with the defaults.  Written after round 1's last GPU minute, simulator-validated only: the bit-plane
#    tests, the lossless float (FPL) encoder tests, the all-integer tile cases.  No -x here: one surprise must not hide the rest.
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider > "$OUT/tests_default.log" 2>&1
#    compute-sanitizer over the FPL encoder (new kernels: PackBits scans, plane Huffman writer)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fpl.py -q -m gpu -p no:cacheprovider -k "encoder and (noisy or sparse or long_runs or depth4 or tiny)" > "$OUT/sanitizer_fpl.log" 2>&1
# 2. parity of the variants (encoder / decoder paths are exercised by the parity, fast-path, fuzz and tile suites)
LERC_B200_ENC=pipe timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fastpath.py tests/test_gpu_tiles.py tests/test_gpu_large.py -x -q -m gpu -p no:cacheprovider > "$OUT/tests_enc_pipe.log" 2>&1
LERC_B200_DEC=closure LERC_B200_DEC_RESOLVE=smem timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fastpath.py tests/test_gpu_fuzz.py tests/test_gpu_large.py -x -q -m gpu -p no:cacheprovider > "$OUT/tests_dec_variants.log" 2>&1
# 3. timing, BASELINE configs[1] (c2) and configs[4] (c5, 4096 tiles)
run_bench c2_default -- --steps 20
run_bench c2_enc_pipe LERC_B200_ENC=pipe -- --steps 20
run_bench c2_dec_closure LERC_B200_DEC=closure -- --steps 20
run_bench c2_dec_resolve_smem LERC_B200_DEC_RESOLVE=smem -- --steps 20
run_bench c2_all LERC_B200_ENC=pipe LERC_B200_DEC=closure LERC_B200_DEC_RESOLVE=smem -- --steps 20
run_bench c5_default -- --workload c5 --strip-rows 4096 --steps 5
run_bench c2l_lossless_float -- --workload c2l --steps 5
run_bench c5_enc_pipe LERC_B200_ENC=pipe -- --workload c5 --strip-rows 4096 --steps 5
for f in "$OUT"/tests_*.log; do echo "== $f"; tail -2 "$f"; done
python - <<'PY'
import glob, json, os
for p in sorted(glob.glob("gpurun_out/r2_first/bench_*.json")):
    try:
        d = json.load(open(p))
        k = {n: round(v["ms_per_step"], 4) for n, v in list(d["roofline"]["kernels"].items())[:6]}
        print(os.path.basename(p), round(d["value"], 2), "Gpx/s", round(d["ms_per_step"], 4), "ms", k)
    except Exception as e:
        print(os.path.basename(p), "ERR", e)
PY
