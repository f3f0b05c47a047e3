#!/usr/bin/env python
"""Where the stall samples of a kernel sit: the stall-reason mix of the first launch whose name matches, and its top
SASS instructions by not-issued samples.
    python tools/ncu_hot.py x.ncu-rep [kernel-name regex] [N]"""
import csv
import subprocess
import sys


def load(rep, kernel):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
    if kernel:
        cmd += ["--kernel-name", "regex:" + kernel]
    rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
    hdr = rows[1]
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":          # next launch
            break
        if len(r) == len(hdr) and r != hdr:
            data.append(r)
    return hdr, data


def main():
    rep = sys.argv[1]
    kernel = sys.argv[2] if len(sys.argv) > 2 else ""
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    hdr, data = load(rep, kernel)
    col = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith("stall_") and h.endswith("(Not Issued)")]
    ni = "Warp Stall Sampling (Not-issued Samples)"
    agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in reasons}
    tot = sum(agg.values()) or 1
    print({k[6:-13]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    t_ni = sum(int(r[col[ni]] or 0) for r in data) or 1
    print("not-issued samples", t_ni, "all", sum(int(r[col["Warp Stall Sampling (All Samples)"]] or 0) for r in data))
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][col[ni]] or 0))[:n]
    for i in sorted(idx):
        r = data[i]
        mix = {h[6:-13]: int(r[col[h]] or 0) for h in reasons if int(r[col[h]] or 0)}
        mix = dict(sorted(mix.items(), key=lambda kv: -kv[1])[:3])
        print(f"{i:5d} {r[col['Source']].strip()[:60]:60s} {100 * int(r[col[ni]] or 0) / t_ni:5.2f}% {mix}")


if __name__ == "__main__":
    main()
