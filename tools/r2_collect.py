#!/usr/bin/env python3
"""tools/r2_collect.py <gpurun_out/dir> -- turns the files of one tools/r2_evidence.sh call into the tracked evidence under profiles/:
bench lines, the ncu launch list, per-kernel summaries of the --set full captures, and profiles/dram_traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum per launch, with the commit the capture was made from)."""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1]
tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
P = os.path.join(ROOT, "profiles")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

for name, dst in [("bench_n1.json", f"{tag}_bench_c2.json"), ("bench_reference_arm.json", f"{tag}_bench_c2_reference_arm.json"), ("bench_c3.json", f"{tag}_bench_c3.json"),
                  ("bench_c4.json", f"{tag}_bench_c4.json"), ("bench_c5.json", f"{tag}_bench_c5.json"), ("bench_c2m.json", f"{tag}_bench_c2m.json"), ("launches.csv", f"{tag}_launches.csv"),
                  ("gpu_tests.log", f"{tag}_gpu_tests.log")]:
    f = os.path.join(src, name)
    if os.path.exists(f):
        shutil.copy(f, os.path.join(P, dst))
commit = open(os.path.join(src, "commit.txt")).read().strip() if os.path.exists(os.path.join(src, "commit.txt")) else "unknown"

traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel, bytes, from the ncu --set full captures of tools/r2_evidence.sh "
                       "(same gpurun call as the bench lines next to this file); workloads as in bench.py", "commit": commit}
for rep, wl in [("prof_c2", "c2"), ("prof_c5", "c5"), ("prof_c4", "c4")]:
    f = os.path.join(src, rep + "_raw.csv")                       # exported on the GPU box by tools/r2_evidence.sh
    if not os.path.exists(f) or os.path.getsize(f) < 100:
        continue
    raw = os.path.join(P, f"{tag}_{wl}_raw.csv")
    shutil.copy(f, raw)
    if os.path.exists(os.path.join(src, rep + "_details.txt")):
        shutil.copy(os.path.join(src, rep + "_details.txt"), os.path.join(P, f"{tag}_{wl}_details.txt"))
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    per = {}
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?")
        short = name.split("(")[0].replace("void ", "").split("<")[0]
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(d[k].replace(",", "")) * UNIT.get(u[k], 1)
        targs = name.split("(")[0]
        key = short + ("_write" if "huffman_runs" in short and (targs.rstrip().endswith("true>") or targs.rstrip().endswith("1>")) else "")
        per.setdefault(key, []).append(tot)
    traffic[wl] = {k: int(max(v)) for k, v in per.items()}          # (the largest launch: strips / warm-up launches are smaller)
    with open(os.path.join(P, f"{tag}_{wl}_summary.txt"), "w") as out:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw], stdout=out, check=False)
json.dump(traffic, open(os.path.join(P, "dram_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
# SASS evidence of the copy-engine (TMA) loads and mbarriers in the headline encoder
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "lerc_b200", "libLerc.so.4")], capture_output=True, text=True).stdout.splitlines()
out, fn = [], None
for line in sass:
    if "Function :" in line:
        fn = line.strip()
    if any(m in line for m in ("UBLKCP", "SYNCS", "ATOMS.OR", "REDUX")) and fn and ("k_encode_tileIfLi3ELb0" in fn or "k_decode_streamIf" in fn):      # the float instantiations of the headline kernels
        out.append(f"{fn[:90]:90s} {line.strip()[:110]}")
open(os.path.join(P, f"{tag}_sass_tma_mbarrier.txt"), "w").write("\n".join(out[:400]) + "\n")
print(len(out), "SASS lines with UBLKCP / SYNCS / ATOMS.OR / REDUX")
