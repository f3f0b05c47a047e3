#!/usr/bin/env python
"""Short single-recording tracking run (BASELINE configs[2] shape) for ncu: 12 channels, N_CYC = 8."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from gps_sdr_receiver_b200 import _capi, synth
from gps_sdr_receiver_b200.tracking import TrackBank
from gps_sdr_receiver_b200._capi import EPOCH_OUT

n_ep = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_rec = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # recordings (aliasing the same samples)
_capi.init(0)
tsats = bench.track_sats(7)
ngps = 8 * 2048
distinct = len(sys.argv) > 3 and sys.argv[3] == "distinct"     # every recording its own samples (default: all alias one buffer)
span = n_ep * ngps
if distinct:
    rec = torch.empty(2 * n_rec * span, dtype=torch.uint8, device="cuda")
    for r in range(n_rec):
        for s0 in range(0, span, 4000 * ngps):
            n = min(4000 * ngps, span - s0)
            synth.make_iq_dev(tsats, n // 2048, noise_sigma=0.25, seed=500 + r, start_sample=s0, out=rec[2 * (r * span + s0):2 * (r * span + s0 + n)], device=0)
else:
    rec = torch.empty(2 * span, dtype=torch.uint8, device="cuda")
    for s0 in range(0, span, 4000 * ngps):
        n = min(4000 * ngps, span - s0)
        synth.make_iq_dev(tsats, n // 2048, noise_sigma=0.25, seed=77, start_sample=s0, out=rec[2 * s0:2 * (s0 + n)], device=0)
out = torch.empty((n_ep, 12 * n_rec, EPOCH_OUT.itemsize), dtype=torch.uint8, device="cuda")
bank = TrackBank(8, 12 * n_rec, device=0)
for r in range(n_rec):
    for s in tsats:
        bank.add(s.prn, 50.0 * np.round(s.doppler / 50.0), (int(s.delay) + 1) % 2048, rec=r)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
bank.process_dev(rec, ngps, n_ep, rec_stride=span if distinct else 0, out=out)
e1.record()
torch.cuda.synchronize()
print(n_rec, "recordings: us per epoch", e0.elapsed_time(e1) * 1e3 / n_ep)
bank.close()
