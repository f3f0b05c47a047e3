"""World-size-2 checks of the multi-GPU plumbing on CPU (gloo): partitioning, the gather
of per-recording tuples, and the merge of Doppler-bin shards.  The per-rank "search" is
the oracle on a tiny grid (this is a test: the oracle may be called here)."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRNS = [4, 9, 11, 30]
BINS = [-2500.0 + 500.0 * b for b in range(6)]


def _oracle_best(raw, prns, bins):
    """ACQ_BEST[1, nprn] from the oracle grid over `bins` (first max of z over bins)."""
    from gps_sdr_receiver_b200._capi import ACQ_BEST
    from oracle import gps_oracle as orc
    out = np.zeros((1, len(prns)), dtype=ACQ_BEST)
    data = orc.raw_to_complex(raw)
    cols = [orc.acq_grid(data, prns, f, 0.0, 1, 1, 4, orc.ACQ_MODE_POW) for f in bins]
    for i, p in enumerate(prns):
        z = np.array([c["z"][i, 0] for c in cols])
        b = int(np.argmax(z))
        out[0, i]["prn"], out[0, i]["bin"] = p, b
        for k in ("mx", "peak", "mean", "std", "z", "em1", "ep1", "second"):
            out[0, i]["cell"][k] = cols[b][k][i, 0]
    return out


def _recordings():
    from gps_sdr_receiver_b200 import synth
    sats = [synth.Sat(prn=4, doppler=-1900.0, delay=300.2, amp=0.09), synth.Sat(prn=11, doppler=40.0, delay=1200.7, amp=0.09)]
    return [synth.make_iq(sats, 4, seed=s) for s in (1, 2, 3)]


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:                      # surface the failure instead of a queue timeout
        q.put((rank, repr(e), None))
        raise


def _worker_body(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gps_sdr_receiver_b200 import multi
    recs = _recordings()
    # (1) recordings partitioned across ranks, tuples gathered
    mine = multi.partition(len(recs), world, rank)
    local = np.concatenate([_oracle_best(recs[r], PRNS, BINS) for r in mine], axis=0)
    allrec = multi.gather_recordings(local, len(recs))
    # (2) Doppler bins of one recording partitioned, shards merged
    mb = multi.partition(len(BINS), world, rank)
    shard = _oracle_best(recs[0], PRNS, [BINS[b] for b in mb])
    merged = multi.gather_bin_shards(shard, len(BINS))
    # (3) per-stream results of a batch (configs[4]): ragged per-rank arrays, rank 1 empty on purpose
    from gps_sdr_receiver_b200.batch import STREAM_RESULT, gather_stream_results
    mine_sr = np.zeros(3 if rank == 0 else 0, dtype=STREAM_RESULT)
    mine_sr["rec"], mine_sr["prn"], mine_sr["freq"] = rank, np.arange(mine_sr.size) + 1, 1000.5 * (rank + 1)
    allsr = gather_stream_results(mine_sr)
    assert allsr.size == 3 and list(allsr["prn"]) == [1, 2, 3] and set(allsr["rec"]) == {0} and allsr["freq"][0] == 1000.5
    q.put((rank, allrec.tobytes(), merged.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_balanced_and_complete():
    from gps_sdr_receiver_b200.multi import partition
    for n in (0, 1, 7, 41, 256):
        for w in (1, 2, 3, 8):
            parts = [partition(n, w, r) for r in range(w)]
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_class_major_bin_partition_and_merge(built_lib):
    """multi.partition_bins / merge_bin_lists (the strong-scaling split of BASELINE configs[3]): complete, balanced, few
    base spectra per rank, and the merge picks the largest z with ties to the lowest GLOBAL bin."""
    from gps_sdr_receiver_b200 import multi
    from gps_sdr_receiver_b200._capi import ACQ_BEST
    from gps_sdr_receiver_b200.acquisition import classify_bins
    bins = [-10000.0 + 50.0 * b for b in range(401)]
    base, _, _ = classify_bins(bins)
    for world in (1, 2, 3, 4, 8):
        lists = [multi.partition_bins(bins, world, r) for r in range(world)]
        assert sorted(b for l in lists for b in l) == list(range(401))
        assert max(map(len, lists)) - min(map(len, lists)) <= 1
        assert all(l == sorted(l) for l in lists)
        assert max(len(set(base[l])) for l in lists) <= -(-20 // world) + 1          # 3 base spectra per rank at world 8, not 20
    rng = np.random.default_rng(3)
    z = rng.integers(0, 6, size=(2, 5, 401)).astype(np.float32)                       # many ties on purpose
    lists = [multi.partition_bins(bins, 4, r) for r in range(4)]
    parts = []
    for l in lists:
        p = np.zeros((2, 5), dtype=ACQ_BEST)
        zl = z[:, :, l]
        p["bin"] = np.argmax(zl, axis=2)                                               # first max within the shard
        p["cell"]["z"] = np.take_along_axis(zl, p["bin"][..., None].astype(np.int64), axis=2)[..., 0]
        parts.append(p)
    merged = multi.merge_bin_lists(parts, lists)
    assert np.array_equal(merged["bin"], np.argmax(z, axis=2)) and np.array_equal(merged["cell"]["z"], z.max(axis=2))


@pytest.mark.timeout(300)
def test_two_ranks_equal_one_rank():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for r in res:
        assert r[2] is not None, f"rank {r[0]} failed: {r[1]}"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    recs = _recordings()
    single_all = np.concatenate([_oracle_best(r, PRNS, BINS) for r in recs], axis=0)
    single_bins = _oracle_best(recs[0], PRNS, BINS)
    for rank, allrec, merged in res:
        assert allrec == single_all.tobytes(), f"rank {rank}: gathered recordings differ from the single-rank result"
        assert merged == single_bins.tobytes(), f"rank {rank}: merged Doppler shards differ from the single-rank result"
    # the injected satellites are where they were put
    assert BINS[int(single_bins[0, 0]["bin"])] in (-2000.0, -1500.0) and int(single_bins[0, 0]["cell"]["mx"]) in (300, 301)
