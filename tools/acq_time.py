#!/usr/bin/env python
"""Development harness: time the acquisition kernel pair on the config-2 grid and (optionally) compare the cells
with a saved run -- the check used while the inverse kernel went through its generations (the scalar ones are
bit-identical to each other; the packed / shared-spectra forms differ by rounding, < 2e-5 relative on peaks).
Switches: GPSB200_ACQ_SCALAR=1, GPSB200_ACQ_NOSHARE=1, GPSB200_ACQ_CTAS=1..3:

    python tools/acq_time.py [recs] [steps] [--save ref.npy | --compare ref.npy]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import bench
    from gps_sdr_receiver_b200 import _capi, synth
    from gps_sdr_receiver_b200.acquisition import AcqPlan, GR_ACQ_POW
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    recs = int(args[0]) if args else 256
    steps = int(args[1]) if len(args) > 1 else 5
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
    res = {"recordings": recs, "ms": ms, "cells_per_s": recs * bench.CELLS_PER_REC / (ms * 1e-3),
           "tflops": recs * bench.FLOP_PER_REC / (ms * 1e-3) / 1e12}
    if "--save" in sys.argv:
        np.save(sys.argv[sys.argv.index("--save") + 1], cells)
    if "--compare" in sys.argv:
        ref = np.load(sys.argv[sys.argv.index("--compare") + 1])
        res["mx_equal"] = bool(np.array_equal(ref["mx"], cells["mx"]))
        res["max_rel_peak"] = float(np.max(np.abs(ref["peak"] - cells["peak"]) / ref["peak"]))
        res["bitwise_equal"] = bool(ref.tobytes() == cells.tobytes())
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
