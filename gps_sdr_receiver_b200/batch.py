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

    def _acquire(self, d_raw, nrec: int, rec_samples: int, rec0: int):
        """Steps 1 + 2: one search launch, then the channels of a fresh TrackBank.  Returns (bank, out) with the
        acquisition columns of STREAM_RESULT filled, or (None, out) when nothing was found."""
        best = AcqPlan.best_from_tensor(self.plan.search_dev(d_raw, nrec=nrec, rec_stride=rec_samples))
        chosen = [select_sats(best[r], self.z_min, self.max_sat) for r in range(nrec)]
        n_ch = sum(len(c) for c in chosen)
        out = np.zeros(n_ch, dtype=STREAM_RESULT)
        if n_ch == 0:
            return None, out
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
        return bank, out

    @staticmethod
    def _summarise(rec_t, out: np.ndarray) -> np.ndarray:
        """Step 4 on the device: reduce the gr_epoch_out records [n_ep, n_ch, 448] (uint8, CUDA) to one STREAM_RESULT per
        channel; only the last epoch's records and three small vectors cross PCIe."""
        import torch
        from ._capi import EPOCH_OUT
        n_ep, n_ch = rec_t.shape[0], rec_t.shape[1]
        cp = rec_t.view(torch.float64)[:, :, EPOCH_OUT.fields["code_phase"][1] // 8]          # [n_ep, n_ch]
        valid = cp >= 0
        idx = torch.arange(n_ep, device=rec_t.device, dtype=torch.int64)[:, None].expand(n_ep, n_ch)
        lastv = torch.where(valid, idx, torch.full_like(idx, -1)).max(dim=0).values
        cp_last = torch.where(lastv >= 0, cp.gather(0, lastv.clamp(min=0)[None, :])[0], torch.full_like(cp[0], -1.0))
        n_valid = valid.sum(dim=0)
        last = TrackBank.records_from_tensor(rec_t[n_ep - 1:n_ep])[0]
        out["locked"], out["sweep"] = last["locked"], last["sweep"]
        out["freq"], out["amplitude"], out["corr_q"] = last["freq"], last["amplitude"], last["corr_q"]
        out["n_code_phase"] = n_valid.cpu().numpy()
        out["code_phase"] = cp_last.cpu().numpy()
        return out

    def run_local(self, d_raw, nrec: int, rec_samples: int, rec0: int = 0) -> np.ndarray:
        """d_raw: torch uint8 CUDA tensor holding `nrec` recordings of `rec_samples` samples each
        (I,Q bytes).  Returns STREAM_RESULT[n_channels] for these recordings (`rec` = rec0 + local index)."""
        import torch
        n_ep = rec_samples // self.ngps
        if n_ep < 1 or rec_samples < self.acq_samples:
            raise ValueError("recordings are shorter than one tracking epoch / the acquisition window")
        bank, out = self._acquire(d_raw, nrec, rec_samples, rec0)
        if bank is None:
            return out
        try:
            rec_t = bank.process_dev(d_raw, self.ngps, n_ep, rec_stride=rec_samples)
            out = self._summarise(rec_t, out)
            torch.cuda.synchronize()
        finally:
            bank.close()
        return out

    def run_host(self, h_raw, nrec: int, rec_samples: int, rec0: int = 0, chunks: int = 8) -> np.ndarray:
        """The same for recordings in (pinned) HOST memory: `h_raw` is a torch uint8 CPU tensor [nrec * 2 * rec_samples].
        The recordings go to the device in `chunks` slices along time on a copy stream; the search runs as soon as the
        first slice is there and every slice is tracked while the next one is still on the bus (state stays in the
        bank between launches), so the upload hides behind the kernels -- or the other way round."""
        import torch
        n_ep = rec_samples // self.ngps
        if n_ep < 1 or rec_samples < self.acq_samples:
            raise ValueError("recordings are shorter than one tracking epoch / the acquisition window")
        dev = torch.device(f"cuda:{self.device}")
        d_raw = torch.empty(nrec * 2 * rec_samples, dtype=torch.uint8, device=dev)
        h2, d2 = h_raw.view(nrec, 2 * rec_samples), d_raw.view(nrec, 2 * rec_samples)
        acq_ep = -(-self.acq_samples // self.ngps)
        per = max(acq_ep, -(-n_ep // max(1, chunks)))
        cuts = list(range(0, n_ep, per)) + [n_ep]
        copy_s, main_s = torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
        events = []
        with torch.cuda.stream(copy_s):
            for a, b in zip(cuts[:-1], cuts[1:]):
                lo, hi = 2 * a * self.ngps, 2 * (b * self.ngps if b < n_ep else rec_samples)
                for r in range(nrec):
                    d2[r, lo:hi].copy_(h2[r, lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_s)
                events.append(ev)
        main_s.wait_event(events[0])
        bank, out = self._acquire(d_raw, nrec, rec_samples, rec0)
        if bank is None:
            torch.cuda.synchronize()
            return out
        try:
            from ._capi import EPOCH_OUT
            rec_t = torch.empty((n_ep, bank.num_active, EPOCH_OUT.itemsize), dtype=torch.uint8, device=dev)
            for (a, b), ev in zip(zip(cuts[:-1], cuts[1:]), events):
                main_s.wait_event(ev)
                bank.process_dev(d_raw[2 * a * self.ngps:], (a + 1) * self.ngps, b - a, rec_stride=rec_samples, out=rec_t[a:b])
            out = self._summarise(rec_t, out)
            torch.cuda.synchronize()
        finally:
            bank.close()
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
