#!/usr/bin/env python
"""Throughput of the GPS L1 C/A hot path on B200 (BASELINE.json metric):
acquisition cells/s (PRN x Doppler x code phase) and tracking x-realtime.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the CPU implementation, same metric

N > 1 is launched by torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.

Workload at every N (weak scaling): BASELINE.json configs[1] "cold-start acquisition":
32 PRN x 41 Doppler bins (+-10 kHz, 500 Hz) x 2048 code phases, 1 ms coherent x 10
non-coherent, synthetic uint8 I/Q at 2.048 MS/s.  One grid is only 2.7 Mcells / 1.8 GFLOP, so a
"step" searches a batch of `--recs` independent 10-ms recordings per GPU in one launch.
The `tracking` object of the same line is configs[2]: 12 channels, 8-ms epochs, a 10-minute
synthetic recording per GPU; `acq_fine` is configs[3] (10 ms x 20, 401 bins) with every rank searching
its own recordings and, for N > 1, `acq_fine_sharded` the same search for ONE set of recordings with
its Doppler bins sharded across the ranks; `tracking_batch` the per-GPU share of configs[4].
See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "acq cells/s (PRN x Doppler x code-phase)"
WORKLOAD = ("cold-start acquisition 32 PRN x 41 Doppler (+-10 kHz, 500 Hz) x 2048 code phases, 1 ms coherent x 10 "
            "non-coherent, uint8 IQ 2.048 MS/s (BASELINE configs[1])")
NPRN, NBIN, NLAG, TCOH, NNONCOH = 32, 41, 2048, 1, 10
BINS = [-10000.0 + 500.0 * b for b in range(NBIN)]
PRNS = list(range(1, NPRN + 1))
CELLS_PER_REC = NPRN * NBIN * NLAG
REC_SAMPLES = TCOH * NNONCOH * 2048
# algorithmic FP32 flops per cell, SURVEY.md 8(d): D*K*F + P*D*K*F + 6*P*D*K*N + 4*P*D*K*N + 8*D*K*Tc*N, F = 5 N log2 N
F_FFT = 5 * 2048 * 11
FLOP_PER_REC = (NBIN * NNONCOH * F_FFT + NPRN * NBIN * NNONCOH * F_FFT + NPRN * NBIN * NNONCOH * 2048 * 10
                + NBIN * NNONCOH * TCOH * 2048 * 8)
FLOP_PER_CELL = FLOP_PER_REC / CELLS_PER_REC          # 669.7
FP32_PEAK_THEORY = 148 * 128 * 2 * 1.965e9 / 1e12     # 74.4 TFLOP/s at clocks.max.sm

TRACK_NCH, TRACK_NCYC = 12, 8


def bench_sats(seed: int, nsat: int = 8):
    from gps_sdr_receiver_b200 import synth
    rng = np.random.default_rng(seed)
    prns = sorted(int(p) for p in rng.permutation(np.arange(1, 33))[:nsat])
    return [synth.Sat(prn=p, doppler=float(np.round(rng.uniform(-9500, 9500), 1)), delay=float(np.round(rng.uniform(2, 2040), 2)),
                      amp=0.09, phi0=float(np.round(rng.uniform(-3, 3), 2)), bit_offset_ms=int(rng.integers(0, 20)), bit_seed=k)
            for k, p in enumerate(prns)]


def track_sats(seed: int = 7):
    from gps_sdr_receiver_b200 import synth
    rng = np.random.default_rng(seed)
    prns = sorted(int(p) for p in rng.permutation(np.arange(1, 33))[:TRACK_NCH])
    return [synth.Sat(prn=p, doppler=float(np.round(rng.uniform(-4200, 4200), 1)), delay=float(np.round(rng.uniform(2, 2040), 2)),
                      amp=0.08, phi0=float(np.round(rng.uniform(-3, 3), 2)), doppler_rate=float(np.round(rng.uniform(-0.5, 0.5), 2)),
                      bit_offset_ms=int(rng.integers(0, 20)), bit_seed=k) for k, p in enumerate(prns)]


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        sm, mx, reasons = [], 0.0, set()
        for ts, line in self.rows:
            if not any(a <= ts <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- CPU legs (the oracle: test infrastructure, used here only as the timed CPU baseline) -------------
def _cpu_make_input(seed):
    from gps_sdr_receiver_b200 import synth
    return synth.make_iq(bench_sats(seed), NNONCOH * TCOH, seed=seed)


def _cpu_grid_worker(raw):
    """One full configs[1] grid on one core; `raw` = the recording's uint8 I/Q (synthesised outside the timed region)."""
    from oracle import gps_oracle as orc
    t0 = time.perf_counter()
    g = orc.acq_grid(orc.raw_to_complex(raw), PRNS, BINS[0], 500.0, NBIN, TCOH, NNONCOH, orc.ACQ_MODE_POW)
    return time.perf_counter() - t0, float(g["z"].max())


FORM_TEXT = {
    "exact": "exact (chosen by gr_acq_plan_create for this grid: largest phase argument 1.26e4 rad): one forward spectrum per bin, the "
             "reference's float32 argument fl32(w32 * fl32((n+1)/fs)) for every sample of every bin (gpsrecv.py:232-235)",
    "fast": "fast: one forward spectrum per 1-kHz class of bins, blocks 1..9 of a coherent interval rotated by one "
            "exp(-i w 2048 b / fs) per block (block 0 with the reference's float32 argument per sample)",
}
TRACK_FORM_TEXT = {
    "exact": "exact (default): every sample rotated by the reference's own float32 phase argument fl32(PHASE + fl32(w * SEC_TIME[n])) "
             "(gpslib.py:1343-1346), once, in a fused fold + prompt pass",
    "fast": "fast: factorised NCO (2 sin/cos per thread and epoch), two sample passes",
}
FINE_BINS = [-10000.0 + 50.0 * b for b in range(401)]
FINE_TCOH, FINE_K = 10, 20
_FINE_RAW = None


def _cpu_fine_init(raw):
    global _FINE_RAW
    from oracle import gps_oracle as orc
    _FINE_RAW = orc.raw_to_complex(raw)
    orc.code_spectrum(1)


def _cpu_fine_worker(bin_idx):
    """The Doppler bins `bin_idx` of ONE configs[3] grid (32 PRN x 2048 lags each) on one core."""
    from oracle import gps_oracle as orc
    t0 = time.perf_counter()
    zmax = 0.0
    for b in bin_idx:
        g = orc.acq_grid(_FINE_RAW, PRNS, FINE_BINS[b], 0.0, 1, FINE_TCOH, FINE_K, orc.ACQ_MODE_POW)
        zmax = max(zmax, float(g["z"].max()))
    return time.perf_counter() - t0, zmax


def cpu_fine_baseline(raw: np.ndarray, min_seconds: float = 8.0):
    """configs[3] on the host: ONE 200-ms recording, its 401 Doppler bins spread over all cores (the grid is what
    shards; the reference itself searches in a single process)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    jobs = [list(range(w, len(FINE_BINS), workers)) for w in range(workers)]
    with mp.get_context("spawn").Pool(workers, initializer=_cpu_fine_init, initargs=(raw,)) as pool:
        pool.map(_cpu_fine_worker, [[0]] * workers)                # warm the workers
        n, spent = 0, 0.0
        while spent < min_seconds:
            t0 = time.perf_counter()
            pool.map(_cpu_fine_worker, jobs)
            spent += time.perf_counter() - t0
            n += 1
    cells = NPRN * len(FINE_BINS) * NLAG
    return {"value": n * cells / spent, "unit": "cells/s", "cores": workers, "kind": "port",
            "sample": f"{n} full grid(s) (32 x 401 x 2048, 10 ms x 20) of one recording, bins interleaved over {workers} worker "
                      f"processes, oracle/gps_oracle.acq_grid, {spent:.1f} s"}


def cpu_acq_baseline(min_seconds: float = 10.0):
    """Single process, like the reference runs acquisition (gpsrecv.py:468-490)."""
    from oracle import gps_oracle as orc
    orc.code_spectrum(1)
    n, spent = 0, 0.0
    while spent < min_seconds:
        dt, _ = _cpu_grid_worker(_cpu_make_input(100 + n))
        spent += dt
        n += 1
    return {"value": n * CELLS_PER_REC / spent, "unit": "cells/s", "cores": 1, "kind": "port",
            "sample": f"{n} full grids (32x41x2048, 1 ms x 10), oracle/gps_oracle.acq_grid (numpy/scipy.fft restatement of "
                      f"gpsrecv.py:217-258 + non-coherent sum), {spent:.1f} s"}


def _cpu_track_worker(job):
    """One worker process = the channels the reference would give to one pool process (gpsrecv.py:300-334)."""
    idx, seconds, repeats = job
    from gps_sdr_receiver_b200 import synth
    from oracle import gps_oracle as orc
    sats = track_sats(7)
    n_ep = int(seconds * 1000) // TRACK_NCYC
    ngps = TRACK_NCYC * 2048
    raw = synth.make_iq(sats, n_ep * TRACK_NCYC, seed=5)          # every worker sees the whole stream, like the pool's workers
    chans = [orc.Channel(sats[i].prn, 50.0 * np.round(sats[i].doppler / 50.0), delay=(int(sats[i].delay) + 1) % 2048, n_cyc=TRACK_NCYC)
             for i in idx]
    t0 = time.perf_counter()
    for rep in range(repeats):                                   # the same samples again: timing only, stream numbers go on
        for e in range(n_ep):
            data = orc.raw_to_complex(raw[e * 2 * ngps:(e + 1) * 2 * ngps])
            for ch in chans:
                ch.process(data, np.int64((rep * n_ep + e + 1) * ngps))
    return time.perf_counter() - t0


def cpu_track_baseline(sats, seconds: float = 2.0, repeats: int = 10):
    """12 channels spread over one worker process per host core (at most one per channel), the reference's pool
    architecture (gpsrecv.py:340-417) with the oracle's restatement of SatStream.process in the workers."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, len(sats)))
    jobs = [([i for i in range(len(sats)) if i % workers == w], seconds, repeats) for w in range(workers)]
    with mp.get_context("spawn").Pool(workers) as pool:
        pool.map(_cpu_track_worker, [([0], 0.1, 1)] * workers)   # warm the workers (imports, code spectra)
        dts = pool.map(_cpu_track_worker, jobs)
    dt = max(dts)
    n_ep = repeats * (int(seconds * 1000) // TRACK_NCYC)
    return {"value": n_ep * TRACK_NCYC * 1e-3 / dt, "unit": "x-realtime", "cores": workers, "kind": "port",
            "sample": f"{n_ep} epochs ({seconds:.0f} s of samples x {repeats}) x {len(sats)} channels (N_CYC=8), one worker process per core ({workers}), each running "
                      f"oracle.Channel.process (restatement of gpslib.SatStream.process) for its channels; slowest worker {dt:.1f} s"}


def run_reference(args):
    """`--impl reference`: the CPU implementation of the same workload on all host cores.  The
    reference is pure Python (numpy + scipy.fft); its own sweepAllSats cannot express the
    1 ms x 10 non-coherent grid, so the timed code is the oracle's restatement of the reference's
    primitives (kind = "port"), one full grid per worker process and step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        # inputs are synthesised (in the workers) BEFORE the timed region: a step times the search only
        warm = [pool.map(_cpu_make_input, range(1000 + w * workers, 1000 + (w + 1) * workers)) for w in range(args.warmup)]
        inputs = [pool.map(_cpu_make_input, range(2000 + k * workers, 2000 + (k + 1) * workers)) for k in range(args.steps)]
        for w in warm:
            pool.map(_cpu_grid_worker, w)
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_cpu_grid_worker, inputs[k])
        dt = time.perf_counter() - t0
    value = args.steps * workers * CELLS_PER_REC / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64/c64 (numpy)", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "recordings_per_step": workers},
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": workers, "kind": "port",
                             "sample": f"{workers} full grids per step, one per worker process (multiprocessing spawn pool)"},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- the CUDA path ----------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from gps_sdr_receiver_b200 import _build, _capi, synth
    from gps_sdr_receiver_b200.acquisition import ACQ_BEST, AcqPlan, GR_ACQ_POW
    from gps_sdr_receiver_b200.tracking import TrackBank
    from gps_sdr_receiver_b200._capi import EPOCH_OUT
    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    _capi.init(local)
    import ctypes as C

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local)
    clocks.start()
    windows = []

    # FP32 FFMA peak of this GPU (roofline denominator; MEASURED_PEAKS.json has HBM and bf16 only)
    tf = C.c_double()
    _capi.check(_capi.lib().gr_debug_fp32_peak(1 << 16, C.byref(tf)))
    fp32_peak = tf.value

    # ---------------- acquisition (configs[1]) ----------------
    R, NBUF = args.recs, 8
    sats = bench_sats(11)          # the same satellites on every rank (the validity checks below were exercised on them); noise differs per rank
    bufs = [synth.make_iq_dev(sats, R * NNONCOH * TCOH, noise_sigma=0.25, seed=1000 * rank + i, device=local) for i in range(NBUF)]
    plan = AcqPlan(PRNS, BINS, TCOH, NNONCOH, GR_ACQ_POW, device=local)
    best_dev = torch.empty((R, NPRN, ACQ_BEST.itemsize), dtype=torch.uint8, device=dev)
    gathered = torch.empty((world, R, NPRN, ACQ_BEST.itemsize), dtype=torch.uint8, device=dev) if world > 1 else None
    cells_dev = torch.empty((R, NPRN, NBIN, 32), dtype=torch.uint8, device=dev)

    # validity of the workload (untimed): every injected satellite is found where it was put
    best = AcqPlan.best_from_tensor(plan.search_dev(bufs[0], nrec=R, out=best_dev))
    torch.cuda.synchronize()
    for s in sats:
        for r in (0, R - 1):
            b = best[r, s.prn - 1]
            assert abs(BINS[int(b["bin"])] - s.doppler) <= 500.0 and (int(b["cell"]["mx"]) - int(s.delay)) % 2048 in (0, 1) \
                and b["cell"]["z"] > 10, ("acquisition missed an injected satellite", s, b)

    def step(i):
        plan.search_dev(bufs[i % NBUF], nrec=R, out=best_dev)
        if world > 1:                                      # gather the per-GPU peak tuples (the only exchange)
            dist.all_gather_into_tensor(gathered, best_dev)

    for i in range(args.warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * R * CELLS_PER_REC / (ms_step * 1e-3)

    # the two grid kernels alone (acq_fwd_kernel 0.5 % + acq_inv_kernel 99.5 %), CUDA events on their stream
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.perf_counter()
    k0.record()
    for i in range(args.steps):
        plan.run_dev(bufs[i % NBUF], nrec=R, out=cells_dev)
    k1.record()
    torch.cuda.synchronize()
    windows.append((t_a, time.perf_counter()))
    ms_kernel = k0.elapsed_time(k1) / args.steps
    inv_kernel_name = plan.inverse_kernel()        # the launcher picks the form of the inverse kernel per call (include/gps_b200.h)
    achieved_tf = R * FLOP_PER_REC / (ms_kernel * 1e-3) / 1e12

    # end to end through the public host API: pinned host I/Q in, host tuples out, every step
    host_in = [torch.empty(bufs[0].numel(), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for h, d in zip(host_in, bufs[:2]):
        h.copy_(d)
    host_np = [h.numpy() for h in host_in]
    best_host = torch.empty((R, NPRN, ACQ_BEST.itemsize), dtype=torch.uint8).pin_memory()
    best_np = best_host.numpy().view(ACQ_BEST).reshape(R, NPRN)
    for i in range(max(1, args.warmup)):
        plan.search(host_np[i % 2], nrec=R, out=best_np)
    barrier()
    t_a = time.perf_counter()
    for i in range(args.steps):
        plan.search(host_np[i % 2], nrec=R, out=best_np)
    t_e2e = time.perf_counter() - t_a
    barrier()
    windows.append((t_a, time.perf_counter()))
    t_e2e = max_over_ranks(t_e2e)
    e2e_value = world * R * CELLS_PER_REC * args.steps / t_e2e
    assert np.array_equal(best_np["bin"], AcqPlan.best_from_tensor(plan.search_dev(bufs[(args.steps - 1) % 2], nrec=R))["bin"])
    launches = args.steps * plan.launches()   # acq_fwd_kernel, inverse kernel (one or two launches, see gr_acq_run_dev), acq_best_kernel per step
    del bufs, cells_dev, host_in

    line = {
        "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "recordings_per_gpu_per_step": R, "cells_per_recording": CELLS_PER_REC,
                   "l2": f"inputs rotate over {NBUF} batches = {NBUF * R * REC_SAMPLES * 2 / 2**20:.0f} MiB > 126 MiB L2",
                   "multi_gpu": "recordings partitioned across ranks; NCCL all_gather of the per-GPU peak tuples each step"},
        "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": R * REC_SAMPLES * 2,
                "d2h_bytes_per_step": R * NPRN * ACQ_BEST.itemsize, "api": "AcqPlan.search -> gr_acq_search_host (C ABI), pinned host buffers"},
        "gpu_launches": launches,
        "roofline": {"bound": "fp32", "kernel": inv_kernel_name + " (+ acq_fwd_kernel, 0.5 % of the launch pair)", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp32_peak,
                     # dram__bytes_read + dram__bytes_write of the kernel pair, ncu --set full capture of 512 recordings of this
                     # workload (profiles/acq_r02_quad_ncu_summary.md: 21.0 + 277.7 MB forward, 339.0 + 22.6 MB inverse), scaled to R.
                     # Against 21.6 MB of algorithmic bytes: the forward spectra (335 MB per 512 recordings, more than L2) are
                     # written once and read back once; 31 GB/s, 0.5 % of HBM bandwidth.
                     "traffic": R * (21.038336 + 277.677056 + 339.014144 + 22.628864) * 1e6 / 512,
                     "peak_source": f"measured in this run: register-resident FFMA chains on all SMs (gr_debug_fp32_peak); "
                                    f"theoretical at 1965 MHz = {FP32_PEAK_THEORY:.1f}",
                     "traffic_source": "ncu --set full capture of 512 recordings, final kernels (profiles/acq_r02_quad_ncu_summary.md), scaled to R",
                     "flop_per_cell": FLOP_PER_CELL, "ms_per_launch": ms_kernel,
                     # what the kernels execute: 2 forward FFTs + wipe-offs per interval instead of the 41 the count credits
                     "flop_executed": R * (FLOP_PER_REC - (NBIN - 2) * NNONCOH * (F_FFT + TCOH * 2048 * 8)),
                     "achieved_executed": R * (FLOP_PER_REC - (NBIN - 2) * NNONCOH * (F_FFT + TCOH * 2048 * 8)) / (ms_kernel * 1e-3) / 1e12,
                     "note": "BASELINE prescribes the FP32 FFT-flop roofline for acquisition; algorithmic flops = 5 N log2 N per FFT of the "
                             "reference's algorithm (one forward FFT per Doppler bin and interval: 4 % of the count); the kernels run 2 "
                             "forward FFTs per interval and derive the other 39 bins as circular shifts"},
    }

    # ---------------- weak-signal fine acquisition (configs[3]): 10 ms x 20, 50 Hz bins, +-10 kHz ----------------
    if not args.skip_fine:
        from gps_sdr_receiver_b200 import multi
        f_bins = FINE_BINS
        f_tcoh, f_k, f_recs = FINE_TCOH, FINE_K, args.fine_recs
        f_flop = (len(f_bins) * f_k * F_FFT + NPRN * len(f_bins) * f_k * F_FFT + NPRN * len(f_bins) * f_k * 2048 * 10
                  + len(f_bins) * f_k * f_tcoh * 2048 * 8)
        f_cells = NPRN * len(f_bins) * NLAG
        f_rec_bytes = 2 * f_tcoh * f_k * 2048
        fsats = [synth.Sat(prn=s.prn, doppler=s.doppler / 2.0, delay=s.delay, amp=0.02, phi0=s.phi0, bit_offset_ms=s.bit_offset_ms,
                           bit_seed=s.bit_seed) for s in sats]
        fraw = synth.make_iq_dev(fsats, f_recs * f_tcoh * f_k, noise_sigma=0.25, seed=4242 + rank, device=local)
        fbest_dev = torch.empty((f_recs, NPRN, ACQ_BEST.itemsize), dtype=torch.uint8, device=dev)

        def fine_parity(plan):
            """Untimed: a sub-grid of recording 0 of THIS run's data against the CPU oracle (peak / mean / std / z relative,
            arg-max lag exact): the two band edges, the centre and the bins of two injected satellites x those two PRNs + two
            PRNs that are not in the recording."""
            from oracle import gps_oracle as orc
            cells = AcqPlan.cells_from_tensor(plan.run_dev(fraw[:f_rec_bytes], nrec=1))[0]
            data = orc.raw_to_complex(fraw[:f_rec_bytes].cpu().numpy())
            absent = [p for p in PRNS if p not in {s_.prn for s_ in fsats}][:2]
            prn_l = [fsats[0].prn, fsats[1].prn] + absent
            bins_l = sorted({0, 200, 400} | {int(round((s_.doppler + 10000.0) / 50.0)) for s_ in fsats[:2]})
            worst, mx_ok = 0.0, True
            for b in bins_l:
                g = orc.acq_grid(data, prn_l, f_bins[b], 0.0, 1, f_tcoh, f_k, orc.ACQ_MODE_POW)
                for i, p_ in enumerate(prn_l):
                    c = cells[p_ - 1, b]
                    for k_ in ("peak", "mean", "std", "z"):
                        worst = max(worst, abs(float(c[k_]) / float(g[k_][i, 0]) - 1.0))
                    if float(g["z"][i, 0]) > 8:
                        mx_ok = mx_ok and int(c["mx"]) == int(g["mx"][i, 0])
            return {"max_rel_err_vs_oracle": worst, "argmax_equal_on_detections": bool(mx_ok),
                    "checked": f"{len(bins_l)} bins x {len(prn_l)} PRNs x 2048 lags of recording 0 (peak, mean, std, z)"}

        def time_fine(plan):
            fb = AcqPlan.best_from_tensor(plan.search_dev(fraw, nrec=f_recs, out=fbest_dev))
            torch.cuda.synchronize()
            for s_ in fsats:                                       # amp 0.02: far below the 1-ms detection limit
                b = fb[0, s_.prn - 1]
                assert abs(f_bins[int(b["bin"])] - s_.doppler) <= 75.0 and b["cell"]["z"] > 10, ("fine acquisition missed", s_, b)
            for _ in range(2):
                plan.search_dev(fraw, nrec=f_recs, out=fbest_dev)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_a = time.perf_counter()
            g0.record()
            for i in range(args.steps):
                plan.search_dev(fraw, nrec=f_recs, out=fbest_dev)
            g1.record()
            plan.search_launches = plan.launches()              # kernel launches of one device search (gr_acq_last_launches)
            plan.search_kernel = plan.inverse_kernel()
            barrier()
            windows.append((t_a, time.perf_counter()))
            return max_over_ranks(g0.elapsed_time(g1)) / args.steps

        fplan = AcqPlan(PRNS, f_bins, f_tcoh, f_k, GR_ACQ_POW, device=local)
        ms_fine = time_fine(fplan)
        # end to end: 16 recordings of 200 ms in pinned host memory -> tuples in host memory, every step
        fhost = torch.empty(fraw.numel(), dtype=torch.uint8).pin_memory()
        fhost.copy_(fraw)
        fbest_host = torch.empty((f_recs, NPRN, ACQ_BEST.itemsize), dtype=torch.uint8).pin_memory()
        fbest_np = fbest_host.numpy().view(ACQ_BEST).reshape(f_recs, NPRN)
        for _ in range(2):
            fplan.search(fhost.numpy(), nrec=f_recs, out=fbest_np)
        barrier()
        t_a = time.perf_counter()
        for i in range(args.steps):
            fplan.search(fhost.numpy(), nrec=f_recs, out=fbest_np)
        t_fe2e = time.perf_counter() - t_a
        fhost_launches = fplan.launches()                       # one host search = its chunks' launches
        barrier()
        windows.append((t_a, time.perf_counter()))
        t_fe2e = max_over_ranks(t_fe2e)
        assert np.array_equal(fbest_np["bin"], AcqPlan.best_from_tensor(fbest_dev)["bin"])
        line["acq_fine"] = {
            "metric": METRIC, "value": world * f_recs * f_cells / (ms_fine * 1e-3), "unit": "cells/s", "ms_per_step": ms_fine,
            "form": FORM_TEXT[fplan.form],
            "config": {"workload": "weak-signal fine acquisition 32 PRN x 401 Doppler (+-10 kHz, 50 Hz) x 2048 code phases, 10 ms coherent "
                                   "x 20 non-coherent (BASELINE configs[3]); every rank searches its own recordings",
                       "recordings_per_gpu_per_step": f_recs, "cells_per_recording": f_cells},
            "e2e": {"value": world * f_recs * f_cells * args.steps / t_fe2e, "unit": "cells/s", "h2d_bytes_per_step": f_recs * f_rec_bytes,
                    "d2h_bytes_per_step": f_recs * NPRN * ACQ_BEST.itemsize, "api": "AcqPlan.search -> gr_acq_search_host (C ABI), pinned host buffers"},
            "roofline": {"bound": "fp32", "achieved": f_recs * f_flop / (ms_fine * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": f_recs * f_flop / (ms_fine * 1e-3) / 1e12 / fp32_peak, "flop_per_cell": f_flop / f_cells,
                         "kernel": fplan.search_kernel + " (+ acq_fwd_kernel, exact form: 15 % of the launches)",
                         # ncu --set full of one 16-recording search (profiles/acq_r02_fine_ncu_summary.md, taken with the 4-CTA form
                         # of the inverse kernel): forward 0.013 + 2.044 GB, inverse 2.156 + 0.013 GB, best 0.007 GB -- one spectrum per
                         # bin, interval and recording (2.1 GB) written once and read once, against 13 MB of samples
                         "traffic": f_recs * (0.013398 + 2.044471 + 2.155828 + 0.013175 + 0.006576) * 1e9 / 16,
                         "traffic_source": "ncu capture (profiles/acq_r02_fine_ncu_summary.md), scaled by recordings"},
            "parity": fine_parity(fplan) if rank == 0 else None,
        }
        launches += (args.steps + 3) * fplan.search_launches + (args.steps + 2) * fhost_launches
        # the fast form forced on the same grid (shared spectra + block rotations): what the exact form costs, and how far
        # the fast form is from the reference at this |f| T
        os.environ["GPSB200_ACQ_EXACT_NCO"] = "0"
        eplan = AcqPlan(PRNS, f_bins, f_tcoh, f_k, GR_ACQ_POW, device=local)
        del os.environ["GPSB200_ACQ_EXACT_NCO"]
        ms_fast = time_fine(eplan)
        line["acq_fine_fast"] = {
            "metric": METRIC, "value": world * f_recs * f_cells / (ms_fast * 1e-3), "unit": "cells/s", "ms_per_step": ms_fast,
            "form": "GPSB200_ACQ_EXACT_NCO=0 (forced): " + FORM_TEXT[eplan.form],
            "roofline": {"bound": "fp32", "achieved": f_recs * f_flop / (ms_fast * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": f_recs * f_flop / (ms_fast * 1e-3) / 1e12 / fp32_peak},
            "parity": fine_parity(eplan) if rank == 0 else None,
        }
        launches += (args.steps + 3) * eplan.search_launches
        eplan.close()
        fplan.close()

        # configs[3] as stated: ONE set of recordings, the Doppler bins of the search sharded across the ranks (strong
        # scaling of one long search), bins grouped by 1-kHz class so that a shard's forward work shrinks with it
        # (multi.partition_bins); NCCL all_gather of the per-shard tuples, merged by largest z / lowest global bin.
        # N = 1 runs the same code (one shard, no gather): the base of the curve.
        lists = [multi.partition_bins(f_bins, world, r) for r in range(world)]
        mine = lists[rank]
        sraw = synth.make_iq_dev(fsats, f_recs * f_tcoh * f_k, noise_sigma=0.25, seed=4242, device=local)   # same bytes on every rank
        splan = AcqPlan(PRNS, [f_bins[b] for b in mine], f_tcoh, f_k, GR_ACQ_POW, device=local)
        # two result buffers: the gather of step i runs on a side stream while step i + 1 is searched (its 20 KB per rank are
        # pure latency); a buffer is searched into again only after its previous gather has completed
        sbest = [torch.empty((f_recs, NPRN, ACQ_BEST.itemsize), dtype=torch.uint8, device=dev) for _ in range(2)]
        sgath = [torch.empty((world, f_recs, NPRN, ACQ_BEST.itemsize), dtype=torch.uint8, device=dev) for _ in range(2)]
        side = torch.cuda.Stream(dev)
        main_s = torch.cuda.current_stream(dev)
        g_done = [None, None]
        s_count = [0]

        def sstep():
            b_ = s_count[0] & 1
            s_count[0] += 1
            if g_done[b_] is not None:
                main_s.wait_event(g_done[b_])
            splan.search_dev(sraw, nrec=f_recs, out=sbest[b_])
            if world > 1:
                ev_ = torch.cuda.Event()
                ev_.record(main_s)
                side.wait_event(ev_)
                with torch.cuda.stream(side):
                    dist.all_gather_into_tensor(sgath[b_], sbest[b_])
                    g_done[b_] = torch.cuda.Event()
                    g_done[b_].record(side)
            else:
                sgath[b_][0].copy_(sbest[b_])

        for _ in range(args.warmup):
            sstep()
        torch.cuda.synchronize()
        last_b = (s_count[0] - 1) & 1
        parts = [AcqPlan.best_from_tensor(sgath[last_b][r]) for r in range(world)]
        merged = multi.merge_bin_lists(parts, lists)
        for s_ in fsats:
            b = merged[0, s_.prn - 1]
            assert abs(f_bins[int(b["bin"])] - s_.doppler) <= 75.0 and b["cell"]["z"] > 10, ("sharded fine acquisition missed", s_, b)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_a = time.perf_counter()
        h0.record()
        for i in range(args.steps):
            sstep()
        for ev_ in g_done:                                 # the timed region ends when the last gathers have completed
            if ev_ is not None:
                main_s.wait_event(ev_)
        h1.record()
        barrier()
        windows.append((t_a, time.perf_counter()))
        ms_sh = max_over_ranks(h0.elapsed_time(h1)) / args.steps
        line["acq_fine_sharded"] = {
            "metric": METRIC, "value": f_recs * f_cells / (ms_sh * 1e-3), "unit": "cells/s", "ms_per_step": ms_sh, "scaling": "strong",
            "config": {"workload": "the same fine grid for ONE set of recordings, its 401 Doppler bins sharded across the ranks by 1-kHz "
                                   "class (BASELINE configs[3]); NCCL all_gather of the per-shard tuples each step (on a side stream, "
                                   "overlapping the next step's search; all gathers complete inside the timed region), merged on every rank",
                       "recordings_per_step": f_recs, "bins_per_rank": len(mine), "cells_per_recording": f_cells},
        }
        launches += (args.steps + args.warmup) * splan.launches()
        splan.close()
        del sraw
        line["gpu_launches"] = launches
        fine_raw_host = fraw[:f_rec_bytes].cpu().numpy() if rank == 0 else None
        del fraw, fhost

    # ---------------- tracking (configs[2]) ----------------
    if not args.skip_tracking:
        n_ep = int(args.track_seconds * 1000) // TRACK_NCYC
        ngps = TRACK_NCYC * 2048
        tsats = track_sats(7)       # same constellation on every rank, different noise
        rec = torch.empty(2 * n_ep * ngps, dtype=torch.uint8, device=dev)
        piece = 4000 * ngps
        for s0 in range(0, n_ep * ngps, piece):
            n = min(piece, n_ep * ngps - s0)
            synth.make_iq_dev(tsats, n // 2048, noise_sigma=0.25, seed=77 + rank, start_sample=s0, out=rec[2 * s0:2 * (s0 + n)], device=local)
        out_dev = torch.empty((n_ep, TRACK_NCH, EPOCH_OUT.itemsize), dtype=torch.uint8, device=dev)

        def new_bank():
            bank = TrackBank(TRACK_NCYC, 16, device=local)
            for s in tsats:     # hand-over from a 50-Hz fine acquisition (configs[3]); the reference's loop
                                # false-locks ~70 Hz off when started > 100 Hz away at N_CYC = 8 (oracle shows the same)
                bank.add(s.prn, 50.0 * np.round(s.doppler / 50.0), (int(s.delay) + 1) % 2048)
            return bank

        wb = new_bank()
        track_form = wb.form
        for _ in range(3):
            wb.process_dev(rec, ngps, min(n_ep, 500), out=out_dev[:min(n_ep, 500)])
        wb.close()
        times = []
        for rep in range(args.track_reps):
            bank = new_bank()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_a = time.perf_counter()
            a0.record()
            bank.process_dev(rec, ngps, n_ep, out=out_dev)
            a1.record()
            barrier()
            windows.append((t_a, time.perf_counter()))
            times.append(max_over_ranks(a0.elapsed_time(a1)) * 1e-3)
            if rep + 1 < args.track_reps:
                bank.close()
        t_track = float(np.mean(times))
        recs = TrackBank.records_from_tensor(out_dev[n_ep - 1:n_ep])[0]
        allr = TrackBank.records_from_tensor(out_dev)
        for ci, s in enumerate(tsats):                        # validity: every channel tracked its satellite to the end
            f_true = s.doppler + s.doppler_rate * args.track_seconds
            assert recs[ci]["locked"] == 1 and abs(recs[ci]["freq"] - f_true) < 5.0 and recs[ci]["sweep"] == 0, (s, recs[ci]["freq"])
            cp = allr[-200:, ci]["code_phase"]
            assert (cp >= 0).mean() > 0.7 and abs(np.median(cp[cp >= 0]) - (s.delay + 0.5)) < 1.0, (s, np.median(cp))
        bank.close()
        # end to end: recording in pinned host memory, records back in host memory (stream pipeline)
        host_rec = torch.empty(rec.numel(), dtype=torch.uint8).pin_memory()
        host_rec.copy_(rec)
        host_out = torch.empty((n_ep, TRACK_NCH, EPOCH_OUT.itemsize), dtype=torch.uint8).pin_memory()
        out_np = host_out.numpy().view(EPOCH_OUT).reshape(n_ep, TRACK_NCH)
        bank = new_bank()
        barrier()
        t_a = time.perf_counter()
        bank.process(host_rec.numpy(), ngps, n_ep, out=out_np)
        t_te2e = time.perf_counter() - t_a
        barrier()
        windows.append((t_a, time.perf_counter()))
        t_te2e = max_over_ranks(t_te2e)
        assert np.array_equal(out_np["delay"], allr["delay"]) and np.array_equal(out_np["freq"], allr["freq"])
        tl = bank.launches()
        bank.close()
        raw_bytes = 2 * n_ep * ngps
        rec_bytes = n_ep * TRACK_NCH * EPOCH_OUT.itemsize
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        # per channel-epoch: two passes of one complex MAC per sample (8 flop) + 2 FFT-2048 + spectrum product and |.|
        flop_ce = 2 * 8 * ngps + 2 * F_FFT + 10 * 2048
        line["tracking"] = {
            "metric": "tracking x-realtime", "value": world * args.track_seconds / t_track, "unit": "x-realtime",
            "config": {"workload": f"steady-state tracking: {TRACK_NCH} channels, 8-ms epochs (N_CYC=8), {args.track_seconds:.0f} s synthetic "
                                   "recording per GPU, loop filters on device (BASELINE configs[2])",
                       "epochs": n_ep, "launches_per_recording": 1, "multi_gpu": "replicas only: one recording per rank"},
            "seconds": t_track, "reps": args.track_reps, "form": TRACK_FORM_TEXT[track_form],
            "e2e": {"value": world * args.track_seconds / t_te2e, "unit": "x-realtime", "h2d_bytes": raw_bytes, "d2h_bytes": rec_bytes,
                    "gpu_launches": tl, "api": "TrackBank.process -> gr_track_process_host (3-stream chunk pipeline), pinned host buffers"},
            "roofline": {"bound": "hbm", "kernel": "track_kernel", "achieved": (raw_bytes + rec_bytes) / t_track / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": (raw_bytes + rec_bytes) / t_track / 1e9 / hbm_peak,
                         # ncu capture of 2000 epochs x 12 channels (profiles/track_r01_v5_ncu_summary.md): 66.3 MB read + 3.1 MB
                         # written = the raw stream once (the 12 channels share it through L2) + the records not still in L2
                         "traffic": n_ep * (66.262016 + 3.059456) * 1e6 / 2000,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else "fallback",
                         "fp32_achieved_tflops": n_ep * TRACK_NCH * flop_ce / t_track / 1e12, "fp32_peak_tflops": fp32_peak,
                         "note": "BASELINE prescribes the HBM roofline; one recording is 12 CTAs running 75 000 dependent epochs, "
                                 "i.e. latency-bound (SURVEY.md 8d); algorithmic bytes = raw I/Q once + one record per channel-epoch"},
        }
        launches += args.track_reps
        line["gpu_launches"] = launches
        del rec, out_dev, host_rec, host_out

        # ---- tracking, many recordings per launch (the per-GPU share of configs[4]: independent recordings) ----
        if args.batch_recs > 0:
            RB, secs = args.batch_recs, args.batch_seconds
            nb_ep = int(secs * 1000) // TRACK_NCYC
            span = nb_ep * ngps
            brec = torch.empty(2 * RB * span, dtype=torch.uint8, device=dev)
            for r in range(RB):
                for s0 in range(0, span, piece):
                    n = min(piece, span - s0)
                    synth.make_iq_dev(tsats, n // 2048, noise_sigma=0.25, seed=500 + r + 100 * rank, start_sample=s0,
                                      out=brec[2 * (r * span + s0):2 * (r * span + s0 + n)], device=local)
            bout = torch.empty((nb_ep, RB * TRACK_NCH, EPOCH_OUT.itemsize), dtype=torch.uint8, device=dev)

            def new_batch_bank():
                bank = TrackBank(TRACK_NCYC, RB * TRACK_NCH, device=local)
                for r in range(RB):
                    for s in tsats:
                        bank.add(s.prn, 50.0 * np.round(s.doppler / 50.0), (int(s.delay) + 1) % 2048, rec=r)
                return bank

            wb = new_batch_bank()
            wb.process_dev(brec, ngps, min(nb_ep, 200), rec_stride=span, out=bout[:min(nb_ep, 200)])
            wb.close()
            def time_batch():
                bank = new_batch_bank()
                form = bank.form
                barrier()
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t_a = time.perf_counter()
                b0.record()
                bank.process_dev(brec, ngps, nb_ep, rec_stride=span, out=bout)
                b1.record()
                barrier()
                windows.append((t_a, time.perf_counter()))
                tb = max_over_ranks(b0.elapsed_time(b1)) * 1e-3
                last = TrackBank.records_from_tensor(bout[nb_ep - 1:nb_ep])[0]
                assert int((last["locked"] == 1).sum()) == RB * TRACK_NCH, "a batched channel lost lock"
                bank.close()
                return tb, form

            t_batch, batch_form = time_batch()
            os.environ["GPSB200_TRK_FAST_NCO"] = "1"            # the factorised NCO (not the parity-verified default), for comparison
            t_batch_fast, _ = time_batch()
            del os.environ["GPSB200_TRK_FAST_NCO"]
            b_raw = 2 * RB * span
            b_out = nb_ep * RB * TRACK_NCH * EPOCH_OUT.itemsize
            line["tracking_batch"] = {
                "metric": "tracking x-realtime, aggregate over independent recordings", "value": world * RB * secs / t_batch,
                "unit": "x-realtime", "seconds": t_batch, "form": TRACK_FORM_TEXT[batch_form],
                "fast_form": {"value": world * RB * secs / t_batch_fast, "unit": "x-realtime", "seconds": t_batch_fast,
                              "form": "GPSB200_TRK_FAST_NCO=1 (forced): " + TRACK_FORM_TEXT["fast"]},
                "config": {"workload": f"{RB} recordings x {TRACK_NCH} channels per GPU, {secs:.0f} s each, 8-ms epochs, one launch "
                                       f"({RB * TRACK_NCH} channel CTAs, all resident: 3 per SM in the kernel's dense form; per-GPU share of BASELINE configs[4])"},
                "roofline": {"bound": "hbm", "kernel": "track_kernel", "achieved": (b_raw + b_out) / t_batch / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": (b_raw + b_out) / t_batch / 1e9 / hbm_peak,
                             # dram__bytes_read + dram__bytes_write of the kernel, ncu --set full capture of 400 epochs x 384
                             # channels on 32 distinct recordings (profiles/track_r02_dense_exact_ncu_summary.md: 477.9 + 65.7 MB
                             # against 488 MB of algorithmic bytes), scaled to this launch
                             "traffic": nb_ep * RB * TRACK_NCH * (477.896192 + 65.671936) * 1e6 / (400 * 384),
                             "traffic_source": "ncu capture of the exact dense form (profiles/track_r02_dense_exact_ncu_summary.md), scaled by channel-epochs",
                             "fp32_achieved_tflops": nb_ep * RB * TRACK_NCH * flop_ce / t_batch / 1e12, "fp32_peak_tflops": fp32_peak,
                             "note": "algorithmic bytes = each recording's raw I/Q once + one record per channel-epoch; the kernel is "
                                     "FP32/latency-bound (about 260 flop per byte, DESIGN.md 4.3), so the HBM fraction stays small by construction"},
            }
            line["gpu_launches"] += 2
            del bout

            # ---- configs[4] per GPU: acquisition -> hand-over -> tracking -> NCCL gather of the per-stream results ----
            from gps_sdr_receiver_b200.batch import BatchReceiver, gather_stream_results
            rx = BatchReceiver(n_cyc=TRACK_NCYC, max_sat=TRACK_NCH, device=local)
            rec_samples = span

            def pipeline(host_tensor=None):
                res = rx.run_host(host_tensor, RB, rec_samples, rec0=rank * RB) if host_tensor is not None \
                    else rx.run_local(brec, RB, rec_samples, rec0=rank * RB)
                return gather_stream_results(res, device=dev) if world > 1 else res

            res = pipeline()                                        # warm-up + validity
            want = {s_.prn for s_ in tsats}
            for r_ in range(rank * RB, (rank + 1) * RB):
                got = {int(x["prn"]) for x in res[res["rec"] == r_] if x["locked"] == 1 and x["sweep"] == 0}
                assert got == want, ("batch pipeline lost a satellite", r_, sorted(got), sorted(want))
            def time_pipeline(arg=None, reps=3):
                """median of `reps` runs (host timer, max over ranks each): the call allocates, copies and synchronises on the
                host side, and the boxes are shared -- single runs scatter by a factor of two"""
                ts, r_ = [], None
                for _ in range(reps):
                    barrier()
                    t_a = time.perf_counter()
                    r_ = pipeline(arg)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t_a
                    barrier()
                    windows.append((t_a, time.perf_counter()))
                    ts.append(max_over_ranks(dt))
                return float(np.median(ts)), r_

            t_pipe, res = time_pipeline()
            n_streams = int(res.size)
            hb = torch.empty(brec.numel(), dtype=torch.uint8).pin_memory()
            hb.copy_(brec)
            del brec
            pipeline(hb)
            t_pipe_h, res_h = time_pipeline(hb)
            assert res_h.tobytes() == res.tobytes()
            rx.close()
            line["batch_pipeline"] = {
                "metric": "recordings/s through acquisition + tracking (60-s recordings)", "value": world * RB / t_pipe, "unit": "recordings/s",
                "x_realtime": world * RB * secs / t_pipe, "seconds": t_pipe, "streams": n_streams,
                "config": {"workload": f"BASELINE configs[4] per GPU: {RB} independent {secs:.0f}-s recordings x {TRACK_NCH} satellites, fine cold-start "
                                       "search (201 bins x 32 PRN, 10 ms x 2) -> hand-over -> 8-ms tracking of every found satellite -> per-stream "
                                       "summaries reduced on the device; NCCL all_gather of the summaries inside the timed region (N > 1)",
                           "recordings_per_gpu": RB, "wall_clock": "host timer around the call, synchronised on both sides, max over ranks; median of 3 calls"},
                "e2e": {"value": world * RB / t_pipe_h, "unit": "recordings/s", "x_realtime": world * RB * secs / t_pipe_h, "seconds": t_pipe_h,
                        "h2d_bytes": 2 * RB * span, "d2h_bytes": int(res.nbytes) + RB * NPRN * ACQ_BEST.itemsize,
                        "api": "BatchReceiver.run_host: pinned host recordings uploaded in 8 time slices on a copy stream behind the kernels"},
            }
            line["gpu_launches"] += 4 * (3 + 1) + 4 * (3 + 8)
            del hb

    clocks.stop()
    line["clocks"] = clocks.summary(windows)

    if rank == 0 and world == 1 and not args.skip_cpu:
        line["cpu_baseline"] = cpu_acq_baseline(args.cpu_seconds)
        if "acq_fine" in line:
            line["acq_fine"]["cpu_baseline"] = cpu_fine_baseline(fine_raw_host, 8.0)
        if "tracking" in line:
            line["tracking"]["cpu_baseline"] = cpu_track_baseline(track_sats(7), 2.0)
    elif rank == 0:
        line["cpu_baseline"] = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--recs", type=int, default=512, help="independent 10-ms recordings per GPU per step")
    ap.add_argument("--track-seconds", type=float, default=600.0)
    ap.add_argument("--track-reps", type=int, default=2)
    ap.add_argument("--batch-recs", type=int, default=32, help="recordings per GPU in the batched tracking line (0 = skip); 32 = 256 / 8, BASELINE configs[4]")
    ap.add_argument("--batch-seconds", type=float, default=60.0)
    ap.add_argument("--fine-recs", type=int, default=16, help="recordings per GPU in the weak-signal fine-acquisition line")
    ap.add_argument("--skip-fine", action="store_true")
    ap.add_argument("--skip-tracking", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
