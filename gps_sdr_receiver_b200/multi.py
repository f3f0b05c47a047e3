"""Multi-GPU plumbing for the acquisition path: one process per GPU (torch.distributed,
NCCL over NVLink on the GPUs, gloo in the CPU tests).

The search grid shards with no data-path collective: either
independent recordings are partitioned across ranks, or -- for one long search
(BASELINE configs[3]) -- the Doppler bins are.  The only exchange is an all_gather of
the small per-(recording, PRN) peak tuples, after which every rank holds the full
answer.  Tracking one recording does not shard (time recurrence): replicas only."""
from __future__ import annotations

import numpy as np

from ._capi import ACQ_BEST


def partition(n_items: int, world: int, rank: int) -> range:
    """Contiguous, balanced slice of range(n_items) owned by `rank`."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def _all_gather_bytes(local: np.ndarray, sizes: list[int], device=None) -> list[np.ndarray]:
    """all_gather of variably sized byte buffers (padded to the largest)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    cap = max(sizes)
    buf = torch.zeros(cap, dtype=torch.uint8, device=device)
    buf[:local.nbytes] = torch.from_numpy(local.view(np.uint8).reshape(-1)).to(buf.device)
    outs = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [outs[r].cpu().numpy()[:sizes[r]] for r in range(world)]


def gather_recordings(best_local: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """Recordings partitioned with `partition`: every rank contributes ACQ_BEST[n_local, nprn]
    and receives ACQ_BEST[n_total, nprn] in recording order."""
    import torch.distributed as dist
    world = dist.get_world_size()
    nprn = best_local.shape[1]
    sizes = [len(partition(n_total, world, r)) * nprn * ACQ_BEST.itemsize for r in range(world)]
    parts = _all_gather_bytes(np.ascontiguousarray(best_local), sizes, device)
    return np.concatenate([p.view(ACQ_BEST).reshape(-1, nprn) for p in parts], axis=0)


def merge_bin_shards(parts: list[np.ndarray], bin_offsets: list[int]) -> np.ndarray:
    """Doppler bins partitioned: parts[r] = ACQ_BEST[nrec, nprn] over rank r's bins (bin
    indices local to the shard).  Returns the global best per (recording, PRN): largest z,
    ties to the lowest global bin -- the same answer as one search over all bins."""
    out = parts[0].copy()
    out["bin"] += bin_offsets[0]
    for p, off in zip(parts[1:], bin_offsets[1:]):
        q = p.copy()
        q["bin"] += off
        better = q["cell"]["z"] > out["cell"]["z"]          # strict: earlier (lower) bins win ties
        out[better] = q[better]
    return out


def partition_bins(bin_hz, world: int, rank: int) -> list[int]:
    """Doppler bins of ONE search split across ranks so that the per-rank forward work shrinks with the shard: bins that
    differ by a multiple of fs / 2048 = 1 kHz share one forward spectrum (acquisition.classify_bins), so the bins are
    ordered class by class and cut into `world` balanced runs -- a rank's ~nbins / world bins then touch 3 or 4 base
    spectra instead of all of them (contiguous 50-Hz bins touch all 20).  Returns this rank's GLOBAL bin indices,
    ascending.  A cell does not depend on the other bins of its plan, so any split is bit-identical to the full search."""
    from .acquisition import classify_bins
    base, _, _ = classify_bins(bin_hz)
    order = sorted(range(len(base)), key=lambda b: (int(base[b]), b))
    return sorted(order[i] for i in partition(len(order), world, rank))


def merge_bin_lists(parts: list[np.ndarray], bin_lists: list[list[int]]) -> np.ndarray:
    """parts[r] = ACQ_BEST[nrec, nprn] over the bins bin_lists[r] (local indices).  Global best per (recording, PRN):
    largest z, ties to the lowest GLOBAL bin -- the answer of one search over all bins."""
    out = None
    for p, bl in zip(parts, bin_lists):
        if len(bl) == 0:
            continue
        q = p.copy()
        q["bin"] = np.asarray(bl, dtype=np.int32)[q["bin"]]
        if out is None:
            out = q
            continue
        zq, zo = q["cell"]["z"], out["cell"]["z"]
        better = (zq > zo) | ((zq == zo) & (q["bin"] < out["bin"]))
        out[better] = q[better]
    return out


def gather_bin_shards(best_local: np.ndarray, n_bins_total: int, device=None) -> np.ndarray:
    import torch.distributed as dist
    world = dist.get_world_size()
    sizes = [best_local.nbytes] * world
    parts = [p.view(ACQ_BEST).reshape(best_local.shape) for p in _all_gather_bytes(np.ascontiguousarray(best_local), sizes, device)]
    offs = [partition(n_bins_total, world, r).start for r in range(world)]
    return merge_bin_shards(parts, offs)
