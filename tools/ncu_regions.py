#!/usr/bin/env python
"""Stall samples of a kernel aggregated between consecutive block barriers (BAR.SYNC), with the opcode mix of each region."""
import csv, subprocess, sys, collections
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
reg, start = [], 0
def flush(i0, i1):
    smp = sum(int(r[col["# Samples"]] or 0) for r in data[i0:i1])
    ex = sum(int(r[col["Instructions Executed"]] or 0) for r in data[i0:i1])
    ops = collections.Counter(r[col["Source"]].split()[0 if not r[col["Source"]].strip().startswith('@') else 1].split('.')[0] for r in data[i0:i1] if int(r[col["Instructions Executed"]] or 0) > 0)
    if smp > tot * 0.004:
        print(f"[{i0:5d},{i1:5d}) samples {smp:7d} ({100*smp/tot:5.1f} %)  warp-inst {ex:9d}  {dict(ops.most_common(6))}")
for i, r in enumerate(data):
    if "BAR.SYNC" in r[col["Source"]]:
        flush(start, i + 1); start = i + 1
flush(start, len(data))
print("total samples", tot)
