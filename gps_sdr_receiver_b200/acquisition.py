"""Host side of the acquisition path: a thin mirror of the functions
`src/gpsrecv.py` runs in-process (sweepAllSats :241-274, findCodePhase :217-227),
on top of the fused CUDA kernel reached through the C ABI (include/gps_b200.h).

`AcqPlan` is the batched, generalised form (PRN x Doppler x code phase, tcoh
coherent x nnoncoh non-coherent, many recordings per launch); `sweepAllSats`
keeps the reference's call signature and first-hit semantics."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi, glob
from ._capi import ACQ_BEST, ACQ_CELL, GR_ACQ_ABS, GR_ACQ_POW, GR_IN_CF32, GR_IN_U8IQ


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def classify_bins(bin_hz, share: bool = True):
    """The 1-kHz classes an AcqPlan forms from its Doppler bins (host only, `gr_acq_classify_bins`): bins that differ by
    a multiple of fs / 2048 share one forward FFT and differ by a circular shift of it.  Returns (base index per bin,
    shift in FFT bins per bin, base frequencies)."""
    b = np.ascontiguousarray(bin_hz, dtype=np.float64)
    base = np.zeros(b.size, np.int32)
    shift = np.zeros(b.size, np.int32)
    base_hz = np.zeros(b.size, np.float64)
    n = _capi.lib().gr_acq_classify_bins(b.ctypes.data, b.size, 1 if share else 0, base.ctypes.data, shift.ctypes.data,
                                         base_hz.ctypes.data)
    _capi.check(min(n, 0))
    return base, shift, base_hz[:n]


class AcqPlan:
    """A fixed search grid: `prns` x `bin_hz`, `tcoh_ms` coherent milliseconds
    (gpsrecv.py:250-254), `nnoncoh` non-coherent accumulations.  Bins 1 kHz apart share one
    forward spectrum (see `classify_bins`); a cell does not depend on the other bins of the plan."""

    def __init__(self, prns, bin_hz, tcoh_ms: int, nnoncoh: int = 1, mode: int = GR_ACQ_POW,
                 in_format: int = GR_IN_U8IQ, device: int = 0):
        _capi.init(device)
        self.prns = np.ascontiguousarray(prns, dtype=np.int32)
        self.bin_hz = np.ascontiguousarray(bin_hz, dtype=np.float64)
        self.tcoh_ms, self.nnoncoh, self.mode, self.in_format = int(tcoh_ms), int(nnoncoh), int(mode), int(in_format)
        self.rec_samples = self.tcoh_ms * self.nnoncoh * glob.CODE_SAMPLES
        h = C.c_void_p()
        _capi.check(_capi.lib().gr_acq_plan_create(self.prns.ctypes.data, len(self.prns), self.bin_hz.ctypes.data,
                                                   len(self.bin_hz), self.tcoh_ms, self.nnoncoh, self.mode,
                                                   self.in_format, C.byref(h)))
        self._h = h

    @property
    def form(self) -> str:
        """'fast' (shared spectra, block rotations) or 'exact' (the reference's float32 phase argument for every sample
        of every bin): chosen by gr_acq_plan_create from the grid's largest phase argument (include/gps_b200.h)."""
        return "exact" if _capi.lib().gr_acq_plan_form(self._h) else "fast"

    @property
    def cells_per_recording(self) -> int:
        return len(self.prns) * len(self.bin_hz) * glob.CODE_SAMPLES

    def close(self):
        if getattr(self, "_h", None):
            try:
                _capi.lib().gr_acq_plan_destroy(self._h)
            except TypeError:          # interpreter teardown: the module globals are already gone
                pass
            self._h = None

    __del__ = close

    def _check(self, n_items: int, nrec: int, rec_stride: int):
        per = 2 if self.in_format == GR_IN_U8IQ else 1
        need = ((nrec - 1) * rec_stride + self.rec_samples) * per
        if n_items < need:
            raise ValueError(f"input holds {n_items} items, the grid needs {need}")

    def run(self, samples, nrec: int = 1, rec_stride: int | None = None) -> np.ndarray:
        """Host buffers in, host results out (H2D + kernel + D2H inside the call).
        `samples`: uint8 I,Q bytes or complex64.  Returns ACQ_CELL[nrec, nprn, nbins]."""
        rec_stride = self.rec_samples if rec_stride is None else int(rec_stride)
        want = np.uint8 if self.in_format == GR_IN_U8IQ else np.complex64
        a = np.ascontiguousarray(samples)
        if a.dtype != want:
            raise TypeError(f"plan expects {np.dtype(want)} samples, got {a.dtype}")
        self._check(a.size, nrec, rec_stride)
        out = np.empty((nrec, len(self.prns), len(self.bin_hz)), dtype=ACQ_CELL)
        _capi.check(_capi.lib().gr_acq_run_host(self._h, a.ctypes.data, nrec, rec_stride, out.ctypes.data))
        return out

    def run_host_into(self, h_samples_ptr: int, nrec: int, rec_stride: int, out: np.ndarray) -> None:
        """Same as run() on pre-allocated (e.g. pinned) buffers, no allocation."""
        _capi.check(_capi.lib().gr_acq_run_host(self._h, h_samples_ptr, nrec, rec_stride, out.ctypes.data))

    def run_dev(self, d_samples, nrec: int = 1, rec_stride: int | None = None, out=None, stream=None):
        """Device tensor in (torch uint8 / complex64 on cuda), device tensor out,
        asynchronous on `stream` (default: torch's current stream).  Returns a uint8
        tensor [nrec, nprn, nbins, 32]; `cells_from_tensor` views it as ACQ_CELL."""
        import torch
        rec_stride = self.rec_samples if rec_stride is None else int(rec_stride)
        self._check(d_samples.numel(), nrec, rec_stride)
        if out is None:
            out = torch.empty((nrec, len(self.prns), len(self.bin_hz), ACQ_CELL.itemsize), dtype=torch.uint8,
                              device=d_samples.device)
        s = torch.cuda.current_stream(d_samples.device).cuda_stream if stream is None else stream
        _capi.check(_capi.lib().gr_acq_run_dev(self._h, d_samples.data_ptr(), nrec, rec_stride, out.data_ptr(), s))
        return out

    @staticmethod
    def cells_from_tensor(t) -> np.ndarray:
        a = t.cpu().numpy()
        return a.view(ACQ_CELL).reshape(a.shape[:-1])

    def launches(self) -> int:
        return _capi.lib().gr_acq_last_launches(self._h)

    def inverse_kernel(self) -> str:
        """Name of the inverse kernel the last run launched (the launcher picks per call, include/gps_b200.h)."""
        return ("acq_inv_kernel", "acq_inv_quad_kernel", "acq_inv_quad_kernel + acq_inv_kernel")[_capi.lib().gr_acq_last_inverse_form(self._h)]

    # ---- the search proper: one (PRN, Doppler bin, code phase, metric) tuple per recording and PRN ----
    def search(self, samples, nrec: int = 1, rec_stride: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers in, ACQ_BEST[nrec, nprn] out (H2D + both kernels + D2H inside the call).
        `samples` may be a pinned numpy view; no allocation when `out` is given."""
        rec_stride = self.rec_samples if rec_stride is None else int(rec_stride)
        a = samples if isinstance(samples, np.ndarray) and samples.flags.c_contiguous else np.ascontiguousarray(samples)
        want = np.uint8 if self.in_format == GR_IN_U8IQ else np.complex64
        if a.dtype != want:
            raise TypeError(f"plan expects {np.dtype(want)} samples, got {a.dtype}")
        self._check(a.size, nrec, rec_stride)
        if out is None:
            out = np.empty((nrec, len(self.prns)), dtype=ACQ_BEST)
        _capi.check(_capi.lib().gr_acq_search_host(self._h, a.ctypes.data, nrec, rec_stride, out.ctypes.data))
        return out

    def search_dev(self, d_samples, nrec: int = 1, rec_stride: int | None = None, out=None, stream=None):
        """Device tensor in, uint8 tensor [nrec, nprn, 40] (ACQ_BEST) out, asynchronous."""
        import torch
        rec_stride = self.rec_samples if rec_stride is None else int(rec_stride)
        self._check(d_samples.numel(), nrec, rec_stride)
        if out is None:
            out = torch.empty((nrec, len(self.prns), ACQ_BEST.itemsize), dtype=torch.uint8, device=d_samples.device)
        s = torch.cuda.current_stream(d_samples.device).cuda_stream if stream is None else stream
        _capi.check(_capi.lib().gr_acq_search_dev(self._h, d_samples.data_ptr(), nrec, rec_stride, out.data_ptr(), s))
        return out

    @staticmethod
    def best_from_tensor(t) -> np.ndarray:
        a = t.cpu().numpy()
        return a.view(ACQ_BEST).reshape(a.shape[:-1])


# ---- gpsrecv-compatible functions ----------------------------------------------------

_PLAN_CACHE: dict = {}


def _cached_plan(prns, bins, tcoh, nnoncoh, mode, fmt) -> AcqPlan:
    key = (tuple(prns), tuple(bins), tcoh, nnoncoh, mode, fmt)
    p = _PLAN_CACHE.get(key)
    if p is None:
        if len(_PLAN_CACHE) > 64:
            for old in _PLAN_CACHE.values():          # free the device memory now, not whenever __del__ runs
                old.close()
            _PLAN_CACHE.clear()
        p = _PLAN_CACHE[key] = AcqPlan(prns, bins, tcoh, nnoncoh, mode, fmt)
    return p


def findCodePhase(cell, corr_min=None):
    """gpsrecv.findCodePhase (gpsrecv.py:217-227) on an already reduced cell."""
    corr_min = glob.CORR_MIN if corr_min is None else corr_min
    z = float(cell["z"])
    return (int(cell["mx"]) if z > corr_min else -1), z


def sweepAllSats(data, freq, satLst, satFound, itSweep=2):
    """Drop-in for gpsrecv.sweepAllSats (gpsrecv.py:241-274): same arguments, same
    in-place mutation of satLst / satFound, same return tuple.  `data` is the
    reference's complex64 stream (or the raw uint8 I,Q bytes).  All Doppler bins of
    the call and all PRNs of satLst are searched in ONE kernel launch; the
    reference's first-hit-wins ordering is then replayed over the returned cells."""
    avg = min(glob.SWEEP_CORR_AVG, glob.N_CYC)
    n = avg * glob.CODE_SAMPLES
    # the bins the reference loop would visit (including its wrap-around), :248,267-272
    bins, f, it, ready = [], freq, 0, False
    while f < glob.MAX_FREQ and it < itSweep:
        bins.append(f)
        f += glob.STEP_FREQ
        if f >= glob.MAX_FREQ:
            ready = True
            f -= glob.MAX_FREQ - glob.MIN_FREQ
        it += 1
    if bins and satLst:
        a = np.asarray(data)
        if a.dtype == np.uint8:
            fmt, a = GR_IN_U8IQ, a[:2 * n]
        else:
            fmt, a = GR_IN_CF32, np.ascontiguousarray(a[:n], dtype=np.complex64)
        prns = list(satLst)
        cells = _cached_plan(prns, bins, avg, 1, GR_ACQ_ABS, fmt).run(a)[0]
        alive = [True] * len(prns)
        for b, fb in enumerate(bins):
            for i, prn in enumerate(prns):
                if alive[i]:
                    delay, z = findCodePhase(cells[i, b])
                    if delay > -1:
                        satFound.append((z, prn, fb, delay))
                        alive[i] = False
                        satLst.remove(prn)
    return ready, f, sorted(satFound, reverse=True)
