#!/usr/bin/env python
"""Stress check of the inverse kernel's launch forms: N launches of each form on the same recordings, every launch's full
cell grid compared bit for bit (on the device) with the 4-CTA form's -- a race in the quad form's stage ring or in the
split launch would show up as an intermittent difference.

    python tools/stress_inverse_forms.py [recs] [launches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from gps_sdr_receiver_b200 import _capi, synth
    from gps_sdr_receiver_b200.acquisition import AcqPlan, GR_ACQ_POW
    recs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    _capi.init(0)
    raw = synth.make_iq_dev(bench.bench_sats(11), recs * 10, noise_sigma=0.25, seed=5, device=0)
    plans = {}
    for q in ("0", "1", "2"):
        os.environ["GPSB200_ACQ_QUAD"] = q
        plans[q] = AcqPlan(bench.PRNS, bench.BINS, 1, 10, GR_ACQ_POW, device=0)
    del os.environ["GPSB200_ACQ_QUAD"]
    ref = plans["0"].run_dev(raw, nrec=recs).clone()
    bad = 0
    for q in ("0", "1", "2"):
        out = torch.empty_like(ref)
        for i in range(n):
            out.zero_()
            plans[q].run_dev(raw, nrec=recs, out=out)
            if not torch.equal(out, ref):
                bad += 1
                print("MISMATCH form", q, "launch", i, int((out != ref).sum()), "bytes differ", flush=True)
        print("form", q, plans[q].inverse_kernel(), n, "launches checked", flush=True)
    print("recordings", recs, "mismatching launches", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
