#!/usr/bin/env python
"""Development harness: time the inverse-FFT acquisition kernel variants (GPSB200_ACQ_VARIANT) on the
config-2 grid and check every variant's cells against variant 0.

    python tools/acq_variants.py 0 0x41 0x43 ...        # driver: one subprocess per variant
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(variant: str, recs: int, steps: int):
    import numpy as np
    import torch
    import bench
    from gps_sdr_receiver_b200 import _capi, synth
    from gps_sdr_receiver_b200.acquisition import AcqPlan, GR_ACQ_POW
    _capi.init(0)
    sats = bench.bench_sats(11)
    bufs = [synth.make_iq_dev(sats, recs * 10, noise_sigma=0.25, seed=i, device=0) for i in range(4)]
    plan = AcqPlan(bench.PRNS, bench.BINS, 1, 10, GR_ACQ_POW, device=0)
    out = torch.empty((recs, 32, 41, 32), dtype=torch.uint8, device="cuda")
    for i in range(3):
        plan.run_dev(bufs[i % 4], nrec=recs, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        plan.run_dev(bufs[i % 4], nrec=recs, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    plan.run_dev(bufs[0], nrec=recs, out=out)
    torch.cuda.synchronize()
    cells = AcqPlan.cells_from_tensor(out).copy()
    ref_path = os.path.join(ROOT, "gpurun_out", "variant0_cells.npy")
    res = {"variant": variant, "ms": ms, "cells_per_s": recs * bench.CELLS_PER_REC / (ms * 1e-3),
           "tflops": recs * bench.FLOP_PER_REC / (ms * 1e-3) / 1e12}
    if int(variant, 0) == 0:
        os.makedirs(os.path.dirname(ref_path), exist_ok=True)
        np.save(ref_path, cells)
    elif os.path.exists(ref_path):
        ref = np.load(ref_path)
        res["mx_equal"] = bool(np.array_equal(ref["mx"], cells["mx"]))
        res["max_rel_peak"] = float(np.max(np.abs(ref["peak"] - cells["peak"]) / ref["peak"]))
        res["bitwise_equal"] = bool(ref.tobytes() == cells.tobytes())
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--one":
        one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    else:
        recs = int(os.environ.get("RECS", "256"))
        steps = int(os.environ.get("STEPS", "5"))
        for v in sys.argv[1:]:
            env = dict(os.environ, GPSB200_ACQ_VARIANT=str(int(v, 0)))
            r = subprocess.run([sys.executable, __file__, "--one", v, str(recs), str(steps)], env=env, capture_output=True, text=True,
                               timeout=300)
            print(r.stdout.strip() or ("FAILED " + v + ": " + r.stderr[-800:]), flush=True)
