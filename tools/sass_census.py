#!/usr/bin/env python
"""Opcode census of the built library (cuobjdump -sass): per kernel the instruction count and the counts of the opcodes that
show what the kernels are made of -- packed FP32 (FFMA2 / FADD2 / FMUL2), tensor-memory loads / stores (LDTM / STTM),
TMA bulk copies (UBLKCP), mbarrier operations (SYNCS), cp.async (LDGSTS), MUFU, shared-memory and global accesses, barriers.

    python tools/sass_census.py [lib.so] > profiles/sass_census_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gps_sdr_receiver_b200", "libgpsb200.so")
KEYS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "LDGSTS", "LDS", "STS",
        "LDG", "STG", "LDL", "STL", "BAR", "REDUX", "SHFL", "DFMA", "DADD", "DMUL", "PRMT", "ELECT"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, counts, order = None, {}, []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[name] = collections.Counter()
        order.append(name)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        counts[name][m.group(1)] += 1
        counts[name]["_total"] += 1
print("# SASS opcode census of", os.path.basename(lib), "(static instruction counts, sm_100a)")
for n in order:
    c = counts[n]
    if c["_total"] < 50:
        continue
    print(f"\n{n}\n  instructions {c['_total']}")
    print("  " + "  ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
