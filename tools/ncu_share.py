#!/usr/bin/env python
"""Share of a kernel's warp samples per range of SASS instructions (the phase shares quoted in DESIGN.md 4.2):
    python tools/ncu_share.py x.ncu-rep <kernel-name regex> <first> <last> <step>      # buckets of `step` instructions
    python tools/ncu_share.py x.ncu-rep <kernel-name regex> b0 b1 b2 ...               # explicit boundaries (>= 4 numbers)"""
import sys

from ncu_hot import load


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    nums = [int(x) for x in sys.argv[3:]]
    hdr, data = load(rep, kernel)
    col = {h: i for i, h in enumerate(hdr)}
    a, ni = "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (Not-issued Samples)"
    tot = sum(int(r[col[a]] or 0) for r in data) or 1
    bounds = list(range(nums[0], nums[1], nums[2])) + [nums[1]] if len(nums) == 3 else [0] + nums + [len(data)]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        seg = data[lo:hi]
        if not seg:
            continue
        s = sum(int(r[col[a]] or 0) for r in seg)
        n = sum(int(r[col[ni]] or 0) for r in seg)
        ops = {}
        for r in seg:
            o = [x for x in r[col["Source"]].strip().split() if not x.startswith("@")][0].split(".")[0]
            ops[o] = ops.get(o, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
        print(f"[{lo:5d},{hi:5d}) all {100 * s / tot:5.2f} %  stalled {100 * n / tot:5.2f} %  {top}")


if __name__ == "__main__":
    main()
