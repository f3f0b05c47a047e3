#!/usr/bin/env python
"""Short run of the config-4 search (32 PRN x 401 bins, 10 ms x 20) for ncu: python tools/prof_fine.py [recs] [steps] [shards]
(shards = N: only the bins rank 0 of N ranks would search, multi.partition_bins -- the per-rank launch of the sharded search)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gps_sdr_receiver_b200 import _capi, synth
from gps_sdr_receiver_b200.acquisition import ACQ_BEST, AcqPlan, GR_ACQ_POW

recs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
_capi.init(0)
sats = bench.bench_sats(11)
fsats = [synth.Sat(prn=s.prn, doppler=s.doppler / 2.0, delay=s.delay, amp=0.02, phi0=s.phi0, bit_offset_ms=s.bit_offset_ms, bit_seed=s.bit_seed) for s in sats]
raw = synth.make_iq_dev(fsats, recs * bench.FINE_TCOH * bench.FINE_K, noise_sigma=0.25, seed=4242, device=0)
shards = int(sys.argv[3]) if len(sys.argv) > 3 else 1
bins = bench.FINE_BINS
if shards > 1:
    from gps_sdr_receiver_b200 import multi
    bins = [bench.FINE_BINS[i] for i in multi.partition_bins(bench.FINE_BINS, shards, 0)]
plan = AcqPlan(bench.PRNS, bins, bench.FINE_TCOH, bench.FINE_K, GR_ACQ_POW, device=0)
best = torch.empty((recs, 32, ACQ_BEST.itemsize), dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
plan.search_dev(raw, nrec=recs, out=best)
e0.record()
for _ in range(steps):
    plan.search_dev(raw, nrec=recs, out=best)
e1.record()
torch.cuda.synchronize()
print(plan.form, "form:", recs, "recordings,", len(bins), "bins,", plan.inverse_kernel(), "ms per search", e0.elapsed_time(e1) / steps)
