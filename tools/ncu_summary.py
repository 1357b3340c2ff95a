#!/usr/bin/env python3
"""tools/ncu_summary.py <raw.csv> [kernel-substring] -- prints the stall reasons, pipe utilisation and DRAM bytes of one `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
sel = sys.argv[2] if len(sys.argv) > 2 else ""
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    if sel and sel not in d.get("Kernel Name", ""):
        continue
    print("==", d.get("Kernel Name", "?")[:80], "| duration", d.get("gpu__time_duration.sum"), "| grid", d.get("launch__grid_size"), "| regs", d.get("launch__registers_per_thread"))
    st = []
    for k, v in d.items():
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") or k.startswith("smsp__average_warp_latency_issue_stalled") and k.endswith(".ratio"):
            try:
                st.append((float(v.replace(",", "")), k))
            except ValueError:
                pass
    for v, k in sorted(st, reverse=True)[:10]:
        print(f"   stall {k.split('issue_stalled_')[1].split('_per_')[0]:24s} {v:7.2f}")
    for key in ["dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]:
        for kk in d:
            if kk.startswith(key):
                print("  ", kk, d[kk])
    for kk in d:
        if kk.startswith("sm__inst_executed_pipe_") and kk.endswith("pct_of_peak_sustained_active"):
            try:
                if float(d[kk]) > 3: print("   pipe", kk.split("pipe_")[1].split(".")[0], d[kk])
            except ValueError:
                pass
