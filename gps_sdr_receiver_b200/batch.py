"""Batches of independent recordings through the whole hot path (BASELINE.json configs[4]).

What `gpsrecv.processData` (src/gpsrecv.py:445-548) does for one live stream -- cold-start
search, `getNewSats` (:423-440), `initPoolStreams` (:385-401), then `satCalc` per stream
(:404-417) -- done for R recordings at once:

    1. ONE acquisition launch over all recordings (fine grid: 10 ms coherent, 50 Hz bins, so
       that the hand-over is inside the pull-in range of the reference's first-order PLL),
    2. per recording the strongest <= max_sat PRNs above threshold become channels of one
       device-resident TrackBank,
    3. ONE tracking launch advances every channel of every recording through all epochs,
    4. the per-channel summaries (STREAM_RESULT) are what leaves the GPU; with
       torch.distributed initialised, recordings are partitioned across ranks
       (multi.partition) and the summaries all-gathered -- the only collective.
"""
from __future__ import annotations

import numpy as np

from . import glob, multi
from .acquisition import ACQ_BEST, AcqPlan, GR_ACQ_POW
from .tracking import TrackBank

STREAM_RESULT = np.dtype([
    ("rec", "<i4"), ("prn", "<i4"), ("acq_bin_hz", "<f4"), ("acq_delay", "<i4"), ("acq_z", "<f4"),
    ("locked", "<i4"), ("sweep", "<i4"), ("n_code_phase", "<i4"),      # epochs that delivered a code phase
    ("freq", "<f8"), ("code_phase", "<f8"),                            # FREQ / last valid codePhase at the end
    ("amplitude", "<f4"), ("corr_q", "<f4"),
])


def select_sats(best_row: np.ndarray, z_min: float, max_sat: int) -> list[int]:
    """Indices (into the plan's PRN list) of the satellites to track for one recording: z above
    threshold, strongest first, at most max_sat (gpsrecv.getNewSats, gpsrecv.py:423-440).
    z_min is well above the reference's CORR_MIN = 8: the statistic here is the largest of
    201 bins x 2048 lags of a |.|^2 sum (chi-square tail), whose noise-only maximum sits near z = 12."""
    z = best_row["cell"]["z"]
    idx = [int(i) for i in np.argsort(-z, kind="stable") if z[i] > z_min]
    return idx[:max_sat]


class BatchReceiver:
    def __init__(self, n_cyc: int = 32, max_sat: int = 12, prns=None, fmin: float = glob.MIN_FREQ, fmax: float = glob.MAX_FREQ,
                 fstep: float = 50.0, tcoh_ms: int = 10, nnoncoh: int = 2, z_min: float = 18.0, device: int = 0):
        self.n_cyc, self.max_sat, self.z_min, self.device = int(n_cyc), int(max_sat), float(z_min), device
        self.prns = list(range(1, 33)) if prns is None else [int(p) for p in prns]
        nb = int(round((fmax - fmin) / fstep)) + 1
        self.bins = [fmin + fstep * b for b in range(nb)]
        self.plan = AcqPlan(self.prns, self.bins, tcoh_ms, nnoncoh, GR_ACQ_POW, device=device)
        self.acq_samples = self.plan.rec_samples
        self.ngps = self.n_cyc * glob.CODE_SAMPLES

    def close(self):
        self.plan.close()

    def run_local(self, d_raw, nrec: int, rec_samples: int, rec0: int = 0) -> np.ndarray:
        """d_raw: torch uint8 CUDA tensor holding `nrec` recordings of `rec_samples` samples each
        (I,Q bytes).  Returns STREAM_RESULT[n_channels] for these recordings (`rec` = rec0 + local index)."""
        import torch
        n_ep = rec_samples // self.ngps
        if n_ep < 1 or rec_samples < self.acq_samples:
            raise ValueError("recordings are shorter than one tracking epoch / the acquisition window")
        best = AcqPlan.best_from_tensor(self.plan.search_dev(d_raw, nrec=nrec, rec_stride=rec_samples))
        chosen = [select_sats(best[r], self.z_min, self.max_sat) for r in range(nrec)]
        n_ch = sum(len(c) for c in chosen)
        out = np.zeros(n_ch, dtype=STREAM_RESULT)
        if n_ch == 0:
            return out
        bank = TrackBank(self.n_cyc, n_ch, device=self.device)
        k = 0
        for r in range(nrec):
            for i in chosen[r]:
                b = best[r, i]
                f, d = self.bins[int(b["bin"])], int(b["cell"]["mx"])
                bank.add(self.prns[i], f, d, rec=r)            # slots ascend in this order = output columns
                out[k]["rec"], out[k]["prn"] = rec0 + r, self.prns[i]
                out[k]["acq_bin_hz"], out[k]["acq_delay"], out[k]["acq_z"] = f, d, b["cell"]["z"]
                k += 1
        recs = TrackBank.records_from_tensor(bank.process_dev(d_raw, self.ngps, n_ep, rec_stride=rec_samples))
        torch.cuda.synchronize()
        bank.close()
        last = recs[-1]
        out["locked"], out["sweep"] = last["locked"], last["sweep"]
        out["freq"], out["amplitude"], out["corr_q"] = last["freq"], last["amplitude"], last["corr_q"]
        cp = recs["code_phase"]                                  # [n_ep, n_ch], -1.0 where the epoch had none
        valid = cp >= 0
        out["n_code_phase"] = valid.sum(axis=0)
        lastv = np.where(valid.any(axis=0), n_ep - 1 - np.argmax(valid[::-1], axis=0), 0)
        out["code_phase"] = np.where(valid.any(axis=0), cp[lastv, np.arange(n_ch)], -1.0)
        return out

    def run(self, d_raw_local, n_total: int, rec_samples: int) -> np.ndarray:
        """Multi-GPU form: every rank holds the recordings multi.partition(n_total, world, rank) in
        `d_raw_local`; returns the gathered STREAM_RESULT of ALL recordings on every rank."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return self.run_local(d_raw_local, n_total, rec_samples)
        world, rank = dist.get_world_size(), dist.get_rank()
        mine = multi.partition(n_total, world, rank)
        local = self.run_local(d_raw_local, len(mine), rec_samples, rec0=mine.start) if len(mine) else np.zeros(0, STREAM_RESULT)
        return gather_stream_results(local, device=d_raw_local.device if d_raw_local is not None else None)


def gather_stream_results(local: np.ndarray, device=None) -> np.ndarray:
    """all_gather of variably sized STREAM_RESULT arrays, concatenated in rank (= recording) order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    n = torch.tensor([local.size], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, n)
    sizes = [int(c.item()) * STREAM_RESULT.itemsize for c in counts]
    if max(sizes) == 0:
        return np.zeros(0, STREAM_RESULT)
    parts = multi._all_gather_bytes(np.ascontiguousarray(local), sizes, device)
    return np.concatenate([np.frombuffer(p.tobytes(), dtype=STREAM_RESULT) for p in parts])
