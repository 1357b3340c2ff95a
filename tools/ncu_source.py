#!/usr/bin/env python3
"""tools/ncu_source.py <src.csv> [N] -- top-N SASS instructions of an `ncu --page source --csv` export by warp stall samples,
plus the sample share between consecutive barrier instructions (phases of the kernel)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
def num(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return 0.0
tot = sum(num(r, "# Samples") for r in body) or 1
toti = sum(num(r, "Instructions Executed") for r in body) or 1
print("total samples", tot, "total warp instructions", toti)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
top = sorted(range(len(body)), key=lambda i: -num(body[i], "# Samples"))[:n]
for i in sorted(top):
    r = body[i]
    ss = sorted(((num(r, s), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {100*num(r,'# Samples')/tot:5.1f}%  exec {num(r,'Instructions Executed'):9.0f}  {r[ix['Source']][:70]:70s} {ss[0][1]}={ss[0][0]:.0f} {ss[1][1]}={ss[1][0]:.0f}")
print("-- phases (between BAR.SYNC instructions): samples%, instr%")
acc_s = acc_i = 0; start = 0
for i, r in enumerate(body):
    acc_s += num(r, "# Samples"); acc_i += num(r, "Instructions Executed")
    if "BAR.SYNC" in r[ix["Source"]] or i == len(body) - 1:
        print(f"  [{start:5d}..{i:5d}] samples {100*acc_s/tot:5.1f}%  instr {100*acc_i/toti:5.1f}%")
        acc_s = acc_i = 0; start = i + 1
