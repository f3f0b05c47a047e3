#!/usr/bin/env python
"""BASELINE configs[4] across GPUs: `torchrun --nproc-per-node N tools/batch_multi.py [n_total] [seconds]`.
Every rank synthesises its share of the recordings on its GPU, runs BatchReceiver.run (acquisition ->
tracking -> NCCL all_gather of the per-stream results) and rank 0 prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from gps_sdr_receiver_b200 import multi, synth
    from gps_sdr_receiver_b200.batch import BatchReceiver
    n_cyc = 32
    n_ms = int(seconds * 1000) // n_cyc * n_cyc
    mine = multi.partition(n_total, world, rank)
    rng_prns = lambda r: sorted(int(p) for p in np.random.default_rng(1000 + r).permutation(np.arange(1, 33))[:8])
    truth = {}
    raw = torch.empty(2 * len(mine) * n_ms * 2048, dtype=torch.uint8, device=f"cuda:{local}")
    for i, r in enumerate(mine):
        rng = np.random.default_rng(2000 + r)
        sats = [synth.Sat(prn=p, doppler=float(np.round(rng.uniform(-4500, 4500), 1)), delay=float(np.round(rng.uniform(2, 2040), 2)),
                          amp=0.07, phi0=float(rng.uniform(-3, 3)), bit_offset_ms=int(rng.integers(0, 20)), bit_seed=100 * r + k)
                for k, p in enumerate(rng_prns(r))]
        truth[r] = sats
        synth.make_iq_dev(sats, n_ms, noise_sigma=0.25, seed=r, out=raw[2 * i * n_ms * 2048:2 * (i + 1) * n_ms * 2048], device=local)
    rx = BatchReceiver(n_cyc=n_cyc, max_sat=12, device=local)
    rx.run(raw, n_total, n_ms * 2048)                          # warm-up (plan scratch, bank allocation)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = rx.run(raw, n_total, n_ms * 2048)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    ok = True
    for r, sats in truth.items():                              # every rank checks its own recordings in the gathered result
        got = {int(x["prn"]): x for x in res[res["rec"] == r]}
        ok &= set(got) == {s.prn for s in sats}
        ok &= all(got[s.prn]["locked"] == 1 and abs(got[s.prn]["freq"] - s.doppler) < 3.0 for s in sats if s.prn in got)
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"workload": f"{n_total} recordings x {n_ms / 1000:.2f} s, 8 satellites each, acquisition + tracking + gather",
                          "n_gpus": world, "seconds": dt, "x_realtime_aggregate": n_total * n_ms / 1000 / dt,
                          "channels": int(res.size), "all_recordings_correct": bool(flag.item())}), flush=True)
    rx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
