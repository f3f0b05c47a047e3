"""Host side of the tracking path.

`TrackBank` is the batched form: all channels of one (or many) recordings advance
`n_epochs` epochs per kernel launch, loop state stays on the device
(include/gps_b200.h, gr_track_*).  It replaces the reference's multiprocessing
pool (src/gpsrecv.py:300-417).

`SatStream` keeps the reference's per-channel interface
(src/gpslib.py:1044-1210: constructor arguments, `.process(data, smpTime)`, the
upper-case state attributes) on top of a one-channel bank, so gpsrecv.runProc can
use it unchanged."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi, glob
from ._capi import EPOCH_OUT, GR_IN_CF32, GR_IN_U8IQ


class TrackBank:
    def __init__(self, n_cyc: int | None = None, max_channels: int = 16, in_format: int = GR_IN_U8IQ,
                 corr_avg: int | None = None, sweep_corr_avg: int | None = None, it_sweep: int | None = None,
                 corr_min: float | None = None, device: int = 0):
        _capi.init(device)
        L = _capi.lib()
        cfg = _capi.TrackCfg()
        _capi.check(L.gr_track_default_cfg(C.byref(cfg)))
        cfg.n_cyc = glob.N_CYC if n_cyc is None else n_cyc
        cfg.corr_avg = glob.CORR_AVG if corr_avg is None else corr_avg
        cfg.sweep_corr_avg = glob.SWEEP_CORR_AVG if sweep_corr_avg is None else sweep_corr_avg
        cfg.it_sweep = glob.IT_SWEEP if it_sweep is None else it_sweep
        cfg.corr_min = glob.CORR_MIN if corr_min is None else corr_min
        cfg.min_freq, cfg.max_freq, cfg.step_freq = glob.MIN_FREQ, glob.MAX_FREQ, glob.STEP_FREQ
        cfg.in_format = in_format
        cfg.max_channels = max_channels
        self.cfg = cfg
        self.n_cyc = cfg.n_cyc
        self.ngps = cfg.n_cyc * glob.CODE_SAMPLES
        self.in_format = in_format
        h = C.c_void_p()
        _capi.check(L.gr_track_bank_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.slots: list[int] = []          # active slots, ascending = column order of the output

    def close(self):
        if getattr(self, "_h", None):
            try:
                _capi.lib().gr_track_bank_destroy(self._h)
            except TypeError:          # interpreter teardown: the module globals are already gone
                pass
            self._h = None

    __del__ = close

    def add(self, prn: int, freq: float, delay: int = 0, rec: int = 0) -> int:
        slot = _capi.check(_capi.lib().gr_track_add(self._h, int(rec), int(prn), float(freq), int(delay)))
        self.slots = sorted(self.slots + [slot])
        return slot

    def remove(self, slot: int) -> None:
        _capi.check(_capi.lib().gr_track_remove(self._h, int(slot)))
        self.slots.remove(slot)

    def request_sweep(self, slot: int) -> None:
        _capi.check(_capi.lib().gr_track_request_sweep(self._h, int(slot)))

    @property
    def num_active(self) -> int:
        return _capi.lib().gr_track_num_active(self._h)

    def launches(self) -> int:
        return _capi.lib().gr_track_last_launches(self._h)

    @property
    def form(self) -> str:
        """'exact' (default: the reference's float32 phase argument for every sample, gpslib.py:1343-1346) or 'fast'
        (factorised NCO, GPSB200_TRK_FAST_NCO=1 when the bank was created); see include/gps_b200.h."""
        return "exact" if _capi.lib().gr_track_bank_form(self._h) else "fast"

    def process(self, samples, smp_time: int, n_epochs: int = 1, nrec: int = 1, rec_stride: int | None = None,
                out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers in, host records out.  Returns EPOCH_OUT[n_epochs, n_active]."""
        a = np.ascontiguousarray(samples)
        want = np.uint8 if self.in_format == GR_IN_U8IQ else np.complex64
        if a.dtype != want:
            raise TypeError(f"bank expects {np.dtype(want)} samples, got {a.dtype}")
        span = n_epochs * self.ngps
        rec_stride = span if rec_stride is None else int(rec_stride)
        per = 2 if self.in_format == GR_IN_U8IQ else 1
        need = ((nrec - 1) * rec_stride + span) * per
        if a.size < need:
            raise ValueError(f"input holds {a.size} items, {n_epochs} epochs need {need}")
        if out is None:
            out = np.zeros((n_epochs, self.num_active), dtype=EPOCH_OUT)
        _capi.check(_capi.lib().gr_track_process_host(self._h, a.ctypes.data, rec_stride, nrec, n_epochs,
                                                      int(smp_time), out.ctypes.data))
        return out

    def process_dev(self, d_samples, smp_time: int, n_epochs: int = 1, rec_stride: int | None = None, out=None,
                    stream=None):
        """Device tensor in, device tensor [n_epochs, n_active, sizeof(gr_epoch_out)] (uint8) out;
        asynchronous on `stream` (default: torch's current stream)."""
        import torch
        rec_stride = n_epochs * self.ngps if rec_stride is None else int(rec_stride)
        if out is None:
            out = torch.empty((n_epochs, self.num_active, EPOCH_OUT.itemsize), dtype=torch.uint8, device=d_samples.device)
        s = torch.cuda.current_stream(d_samples.device).cuda_stream if stream is None else stream
        _capi.check(_capi.lib().gr_track_process_dev(self._h, d_samples.data_ptr(), rec_stride, n_epochs, int(smp_time),
                                                     out.data_ptr(), s))
        return out

    @staticmethod
    def records_from_tensor(t) -> np.ndarray:
        a = t.cpu().numpy()
        return a.view(EPOCH_OUT).reshape(a.shape[:-1])


def prompt_values(rec) -> np.ndarray:
    """gpsData of one epoch record as complex64[n_prompt]."""
    n = int(rec["n_prompt"])
    return np.ascontiguousarray(rec["prompt"][:2 * n]).view(np.complex64)


def prompt_sample_times(rec) -> np.ndarray:
    """ST + n0 of every prompt value (gpslib.py:1408-1440)."""
    n = int(rec["n_prompt"])
    st = np.full(n, int(rec["prompt_st0"]), dtype=np.int64)
    if n > 1:
        st[1:] += int(rec["prompt_b1"]) + glob.CODE_SAMPLES * np.arange(n - 1, dtype=np.int64)
    return st


def new_edges(rec) -> list[tuple[int, int]]:
    """The (MS_TIME, sample time) tuples appended to EDGES during this epoch."""
    mask = int(rec["edge_mask"])
    if not mask:
        return []
    n = int(rec["n_prompt"])
    st = prompt_sample_times(rec)
    ms0 = int(rec["ms_time"]) - n
    return [(ms0 + k, int(st[k])) for k in range(n) if (mask >> k) & 1]


class SatStream:
    """Drop-in for gpslib.SatStream's signal path (src/gpslib.py:1044-1446): same
    constructor, same `process(data, smpTime, sweep=False)` return tuple
    `(SWEEP, frameLst, codePhase, (CORR_Q, CORR_L))`, same state attribute names.

    `data` is what gpsrecv hands over: complex64[NGPS] (src/gpsrecv.py:168-173), or
    the raw uint8 I,Q bytes of the same stream.  Navigation-message decoding
    (evalEdges and below) is outside the hot path: `frameLst` carries the
    once-per-second report dict ('SAT','AMP','CRM','FRQ','SWP', gpslib.py:1124-1131)
    and `EDGES` is kept exactly like the reference keeps it, so the reference's
    decoder can be attached with `frame_decoder`."""

    def __init__(self, satNo, freq, itSweep=10, corrMin=8, corrAvg=8, sweepCorrAvg=4, delay=0,
                 in_format: int | None = None, frame_decoder=None, device: int = 0, bank: TrackBank | None = None):
        self.SAT_NO = satNo
        self._args = dict(corr_avg=corrAvg, sweep_corr_avg=sweepCorrAvg, it_sweep=itSweep, corr_min=corrMin, device=device)
        self._init = (int(satNo), float(freq), int(delay))
        self._banks: dict[int, tuple[TrackBank, int]] = {}
        self._fmt = in_format
        self._bank = None
        self._shared = bank is not None            # a slot of somebody else's bank (ChannelPool)
        if bank is not None:
            self._bank, self._fmt = bank, bank.in_format
            self._slot = bank.add(*self._init)
        self._decoder = frame_decoder
        self.NO_SEC = 1024 // glob.N_CYC
        self.EDGES = [0]
        self.PHASE_LOCKED = False
        self.PHASE = 0.0
        self.FREQ = freq
        self.DELAY = delay
        self.MS_TIME = 0
        self.SMP_TIME = 0
        self.STD_DEV = 0.005
        self.AMPLITUDE = 0.0
        self.MAX_CORR = 0.0
        self.SWEEP = False
        self.CORR_Q = 0
        self.CORR_L = 0
        self.REP_SWEEP = False
        self.GPSDATA = []
        self.last = None

    def _ensure_bank(self, fmt: int):
        if self._bank is None:
            self._fmt = fmt
            self._bank = TrackBank(glob.N_CYC, 1, fmt, **self._args)
            self._slot = self._bank.add(*self._init)
        elif fmt != self._fmt:
            raise TypeError("a SatStream must be fed one sample format (uint8 I/Q or complex64) for its whole life")

    def process(self, data, smpTime, sweep=False):
        a = np.asarray(data)
        fmt = GR_IN_U8IQ if a.dtype == np.uint8 else GR_IN_CF32
        if fmt == GR_IN_CF32 and a.dtype != np.complex64:
            a = a.astype(np.complex64)
        self._ensure_bank(fmt)
        if sweep:
            self._bank.request_sweep(self._slot)
        rec = self._bank.process(a, int(smpTime), 1)[0, 0]
        return self._absorb(rec, int(smpTime))

    def _absorb(self, rec, smpTime):
        self.last = rec
        self.SMP_TIME = smpTime
        if rec["erased"] & 1:                        # erasePrevData / initSweep before the epoch
            self.EDGES = [0]
        if rec["tracked"]:
            self.GPSDATA = prompt_values(rec)
            if rec["locked_in"] and rec["n_prompt"] > 0:
                if self.EDGES[0] == 0:               # sign of the first signal, gpslib.py:1419-1421
                    self.EDGES[0] = float(np.sign(rec["prompt"][0]))
                self.EDGES += new_edges(rec)
        else:
            self.GPSDATA = []
        self.SWEEP = bool(rec["sweep"])
        self.FREQ = float(rec["freq"]) if rec["freq_weak"] else np.float32(rec["freq"])
        self.PHASE = np.float32(rec["phase"])
        self.DELAY = int(rec["delay"])
        self.PHASE_LOCKED = bool(rec["locked"])
        self.MS_TIME = int(rec["ms_time"])
        self.STD_DEV = np.float32(rec["std_dev"])
        self.AMPLITUDE = np.float32(rec["amplitude"])
        self.MAX_CORR = float(rec["max_corr"])
        self.CORR_Q = float(rec["corr_q"])
        self.CORR_L = float(rec["corr_l"])
        frameLst = []
        if rec["report"]:
            if rec["tracked"] and rec["locked_in"]:
                frameLst = self._eval_edges()
            if len(frameLst) == 0:
                frameLst = [{}]
            for dct in frameLst:                     # reportValues, gpslib.py:1124-1131
                dct["SAT"] = self.SAT_NO
                dct["AMP"] = self.AMPLITUDE
                dct["CRM"] = self.MAX_CORR
                dct["FRQ"] = float(rec["report_freq"])
                dct["SWP"] = bool(rec["rep_sweep"])
        if rec["erased"] & 2:                        # initSweep at the end of the epoch
            self.EDGES = [0]
        if len(self.EDGES) != int(rec["edge_len"]) or float(self.EDGES[0]) != float(rec["edge0"]):
            raise RuntimeError(f"host EDGES mirror out of step with the device: {self.EDGES[:3]} "
                               f"vs len {int(rec['edge_len'])} sign {int(rec['edge0'])}")
        codePhase = float(rec["code_phase"])
        return self.SWEEP, frameLst, codePhase, (self.CORR_Q, self.CORR_L)

    def _eval_edges(self):
        """evalEdges (gpslib.py:1451-1462): hand EDGES to the attached decoder (the
        reference's logicalBits/evalGpsBits), then trim like logicalBits does."""
        frames = []
        n = len(self.EDGES)
        if n > 2:
            if self._decoder is not None:
                frames = self._decoder(self, list(self.EDGES)) or []
            last = self.EDGES[0] * (-1) ** (n - 2)
            self.EDGES = [last, self.EDGES[-1]]
        return frames

    def absorb(self, rec, smpTime):
        """Update the host mirror from this channel's gr_epoch_out record (batched use)."""
        return self._absorb(rec, int(smpTime))

    def close(self):
        if self._bank is not None:
            if self._shared:
                if getattr(self._bank, "_h", None):
                    self._bank.remove(self._slot)
            else:
                self._bank.close()
            self._bank = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
