"""BASELINE configs[0] stand-in, closing the loop (build container only; RUNS THE REAL REFERENCE CONSUMER).

Input: gpurun_out/fix_inputs.json written on the GPU box by tools/dump_fix_inputs.py (a copy is committed as
tests/golden/e2e_fix_inputs_gpu_and_cpu.json) -- the once-per-second
(skippedData, frameLst, coPhLst) messages of one synthetic recording, produced (a) by the GPU hot path and (b) by
the CPU oracle (bit-exact restatement of the reference's SatStream) from the SAME bytes.
Both streams go through the unmodified `gpseval.prepCodePhase / evalData / ecefPositions` (SatOrbit,
leastSquaresPos ...), exactly as gpseval.processData chains them (src/gpseval.py:529-540).
Output: tests/golden/e2e_gpu_vs_cpu_through_reference_consumer.json -- per-fix distance between the two streams'
positions and their distance from the simulated truth.

    python oracle/e2e_consumer_on_gpu_outputs.py
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference/src")


def fresh_gpseval():
    for name in ("matplotlib", "matplotlib.pyplot", "gpsui", "gpxpy"):
        m = types.ModuleType(name)
        m.use = lambda *a, **k: None
        m.ion = lambda *a, **k: None
        sys.modules.setdefault(name, m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules.pop("gpseval", None)                     # module-level state (ORB_LIST, COPH_LIST ...) must start empty
    return importlib.import_module("gpseval")


def run_consumer(msgs):
    gpseval = fresh_gpseval()
    fixes = []
    for skipped, frame_lst, coph in msgs:
        coPhLst = {int(k): [(int(n), float(c)) for n, c in v] for k, v in coph.items()}
        cpLst, _ = gpseval.prepCodePhase(coPhLst, 0)
        satResLst, _, _, _, _ = gpseval.evalData(frame_lst, cpLst, {}, {})
        _, recPosLst, _ = gpseval.ecefPositions(satResLst, None)
        for satNo in coPhLst:
            gpseval.COPH_LIST[satNo] = gpseval.COPH_LIST.get(satNo, []) + coPhLst[satNo]
        fixes += [[float(v) for v in p] for p in recPosLst]
    return np.array(fixes)


def main():
    src = os.path.join(ROOT, "gpurun_out", "fix_inputs.json")
    if not os.path.exists(src):                          # the copy committed with the result
        src = os.path.join(ROOT, "tests", "golden", "e2e_fix_inputs_gpu_and_cpu.json")
    d = json.load(open(src))
    truth = np.array(d["truth_ecef"])
    fg, fc = run_consumer(d["gpu"]), run_consumer(d["cpu"])
    assert len(fg) == len(fc) and len(fg) > 0, (len(fg), len(fc))
    assert np.array_equal(fg[:, 0], fc[:, 0])            # same fix epochs
    diff = np.linalg.norm(fg[:, 1:4] - fc[:, 1:4], axis=1)
    eg, ec = np.linalg.norm(fg[:, 1:4] - truth, axis=1), np.linalg.norm(fc[:, 1:4] - truth, axis=1)
    out = {"n_fixes": int(len(fg)), "gpu_vs_cpu_max_m": float(diff.max()), "gpu_vs_cpu_median_m": float(np.median(diff)),
           "gpu_fix_err_median_m": float(np.median(eg)), "cpu_fix_err_median_m": float(np.median(ec)),
           "gpu_mean_fix_err_m": float(np.linalg.norm(fg[:, 1:4].mean(axis=0) - truth)),
           "cpu_mean_fix_err_m": float(np.linalg.norm(fc[:, 1:4].mean(axis=0) - truth)),
           "note": "positions from the unmodified gpseval consumer fed with GPU-tracked vs oracle-tracked messages of the same recording"}
    print(json.dumps(out, indent=1))
    with open(os.path.join(ROOT, "tests", "golden", "e2e_gpu_vs_cpu_through_reference_consumer.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
