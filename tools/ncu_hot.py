#!/usr/bin/env python
"""Where the stall samples of a kernel sit: top SASS instructions by not-issued samples, with the reason mix.
    python tools/ncu_hot.py x.ncu-rep [N]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and h.endswith("(Not Issued)")]
data = rows[2:]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(r[col["Warp Stall Sampling (Not-issued Samples)"]] or 0) for r in data)
tot_all = sum(int(r[col["Warp Stall Sampling (All Samples)"]] or 0) for r in data)
print("total samples", tot_all, "not-issued", tot)
agg = {}
for r in data:
    for h in reasons:
        agg[h] = agg.get(h, 0) + int(r[col[h]] or 0)
print({k[6:-13]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["Warp Stall Sampling (Not-issued Samples)"]] or 0))[:N]
for i in sorted(idx):
    r = data[i]
    mix = {h[6:-13]: int(r[col[h]] or 0) for h in reasons if int(r[col[h]] or 0)}
    print(f"{i:5d} {r[col['Source']].strip()[:70]:70s} ni={r[col['Warp Stall Sampling (Not-issued Samples)']]:>6s} {mix}")
