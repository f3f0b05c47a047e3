#!/usr/bin/env python
"""Development: time the configs[4] per-GPU pipeline (BatchReceiver.run_local / run_host) and its pieces."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from gps_sdr_receiver_b200 import _capi, synth
from gps_sdr_receiver_b200.batch import BatchReceiver
from gps_sdr_receiver_b200.tracking import TrackBank

RB = int(sys.argv[1]) if len(sys.argv) > 1 else 32
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
_capi.init(0)
tsats = bench.track_sats(7)
ngps = 8 * 2048
nb_ep = int(secs * 1000) // 8
span = nb_ep * ngps
brec = torch.empty(2 * RB * span, dtype=torch.uint8, device="cuda")
piece = 4000 * ngps
for r in range(RB):
    for s0 in range(0, span, piece):
        n = min(piece, span - s0)
        synth.make_iq_dev(tsats, n // 2048, noise_sigma=0.25, seed=500 + r, start_sample=s0, out=brec[2 * (r * span + s0):2 * (r * span + s0 + n)], device=0)
rx = BatchReceiver(n_cyc=8, max_sat=12, device=0)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bank, out = rx._acquire(brec, RB, span, 0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rec_t = bank.process_dev(brec, ngps, nb_ep, rec_stride=span)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    out = rx._summarise(rec_t, out)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    recs = TrackBank.records_from_tensor(rec_t)
    bank.close()
    t4 = time.perf_counter()
    dl = recs["delay"].astype(np.int64)
    chg = (dl[1:] != dl[:-1]).mean(axis=0)
    print(f"form {bank.form if hasattr(bank, 'form') else '?'} acquire {t1 - t0:.4f} track {t2 - t1:.4f} summarise {t3 - t2:.4f} close {t4 - t3:.4f}; "
          f"DELAY changes per epoch: mean {chg.mean():.3f} max {chg.max():.3f}; sweeps {int((recs['sweep'] != 0).sum())} tracked {float((recs['tracked'] != 0).mean()):.3f}", flush=True)
    del rec_t, recs
hb = torch.empty(brec.numel(), dtype=torch.uint8).pin_memory()
hb.copy_(brec)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = rx.run_host(hb, RB, span)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"run_host {t1 - t0:.4f}", flush=True)
