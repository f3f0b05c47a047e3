"""CPU oracle for the acquisition + tracking hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of the algorithm that
annappo/GPS-SDR-Receiver runs on the CPU for the path named in BASELINE.json
(`north_star`).  It is the checker for the CUDA path: only `tests/`,
`__graft_entry__.smoke()`, `bench.py`'s cpu_baseline / `--impl reference` legs
and `oracle/make_golden.py` may import it.  Nothing under
`gps_sdr_receiver_b200/` imports it and the product path has no CPU fallback.

Parity status: PINNED AGAINST THE REFERENCE ITSELF.  The reference ships no
tests or golden vectors, so `oracle/make_golden.py` runs
the real reference modules from /root/reference/src in the build container on
seeded synthetic recordings, asserts that this restatement reproduces them
BIT-EXACTLY (same numpy expressions, same dtypes, same evaluation order) and
commits the reference's outputs under tests/golden/.  The generalised
acquisition grid (`acq_grid`, non-coherent accumulation, arbitrary bins) has no
counterpart in the reference; it is built from the pinned primitives and is
"reference-derived", see DESIGN.md.

All `file:line` citations are into /root/reference/src/.  The reference is the
code *as it executes under numpy 2.x* (NEP 50 scalar promotion): after the
first PLL update the carrier frequency and phase are np.float32 scalars.
"""
from __future__ import annotations

import numpy as np
from scipy.fft import fft, ifft

# ---- constants (gpsglob.py:63-75, 119-131) ---------------------------------
CODE_SAMPLES = 2048
SAMPLE_RATE = 1000 * CODE_SAMPLES
MIN_FREQ = -5000.0
MAX_FREQ = +5000.0
STEP_FREQ = 200
CORR_AVG = 8
CORR_MIN = 8
SWEEP_CORR_AVG = 4
IT_SWEEP = 40
IT_SWEEP_ALL = 10
F32 = np.float32
C64 = np.complex64

# G2 taps, IS-GPS-200 (the reference stores the resulting chips literally,
# cacodes.py:5-80; make_golden.py checks equality for PRN 1..37).
_G2_TAPS = ((2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10),
            (2, 3), (3, 4), (5, 6), (6, 7), (7, 8), (8, 9), (9, 10), (1, 4), (2, 5),
            (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6), (5, 7), (6, 8), (7, 9),
            (8, 10), (1, 6), (2, 7), (3, 8), (4, 9), (5, 10), (4, 10), (1, 7), (2, 8),
            (4, 10))


def ca_chips(prn: int) -> np.ndarray:
    """cacodes.py:5-80 — 1023 chips, +1 for logical 1, -1 for logical 0."""
    a, b = _G2_TAPS[prn - 1]
    g1 = np.ones(10, dtype=np.int64)
    g2 = np.ones(10, dtype=np.int64)
    chips = np.empty(1023, dtype=np.int8)
    for i in range(1023):
        chips[i] = 1 if (g1[9] ^ g2[a - 1] ^ g2[b - 1]) else -1
        n1 = g1[2] ^ g1[9]
        n2 = g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9]
        g1[1:] = g1[:-1].copy()
        g2[1:] = g2[:-1].copy()
        g1[0] = n1
        g2[0] = n2
    return chips


_CODE_CACHE: dict[int, np.ndarray] = {}
_SPEC_CACHE: dict[int, np.ndarray] = {}


def ca_code_2048(prn: int) -> np.ndarray:
    """gpslib.py:62-77 (doubledCacode + GPSCacode): every chip twice (2046
    float32 values), then linear interpolation onto 2048 points of a float32
    linspace(0, 2045).  Returns float64[2048]."""
    if prn not in _CODE_CACHE:
        y = np.repeat(ca_chips(prn), 2).astype(F32)
        x = np.arange(len(y), dtype=F32)
        xp = np.linspace(x[0], x[-1], CODE_SAMPLES, endpoint=True, dtype=F32)
        _CODE_CACHE[prn] = np.interp(xp, x, y)
    return _CODE_CACHE[prn]


def code_spectrum(prn: int) -> np.ndarray:
    """gpsrecv.py:574-577 / gpslib.py:1065 — fft of the resampled code."""
    if prn not in _SPEC_CACHE:
        _SPEC_CACHE[prn] = fft(ca_code_2048(prn))
    return _SPEC_CACHE[prn]


def raw_to_complex(raw: np.ndarray) -> np.ndarray:
    """gpsrecv.py:168-173 — uint8 interleaved I,Q -> complex64 in [-1, 1]."""
    words = np.ascontiguousarray(raw).view(np.uint16)
    im, re = np.divmod(words, 256)
    return np.asarray(re + 1j * im, dtype=C64) / 127.5 - (1 + 1j)


def sec_time(n: int) -> np.ndarray:
    """gpsrecv.py:32-33 / gpslib.py:1053-1054 — t[k] = (k+1)/fs as float32."""
    return np.linspace(1, n, n, endpoint=True, dtype=F32) / SAMPLE_RATE


def wipeoff(data, freq, phase, n, t):
    """gpsrecv.py:232-235 / gpslib.py:1343-1346 (demodDoppler)."""
    rot = np.exp(-1j * (phase + 2 * np.pi * freq * t[:n]))
    phase += 2 * np.pi * freq * t[n - 1]
    return rot * data[:n], np.remainder(phase, 2 * np.pi)


def peak_test(corr, corr_min=CORR_MIN):
    """gpsrecv.py:217-227 (findCodePhase): first-max argmax, population std,
    strict '>' against corr_min."""
    mean = np.mean(corr)
    std = np.std(corr)
    mx = np.argmax(corr)
    z = (corr[mx] - mean) / std
    return (mx if z > corr_min else -1), z


def coherent_spectrum(data, first_ms, n_ms):
    """gpsrecv.py:250-254 / gpslib.py:1316-1323 — mean of the complex64 FFTs of
    n_ms consecutive 1-ms blocks."""
    acc = 0
    for i in range(first_ms, first_ms + n_ms):
        acc += fft(data[i * CODE_SAMPLES:(i + 1) * CODE_SAMPLES])
    return acc / n_ms


def sweep_all_sats(data, freq, sat_list, found, spectra, it_sweep=2, n_cyc=32, t=None):
    """gpsrecv.py:241-274 (sweepAllSats).  `spectra[prn]` = code_spectrum(prn).
    Mutates sat_list / found like the reference."""
    if t is None:
        t = sec_time(n_cyc * CODE_SAMPLES)
    ready = False
    avg = min(SWEEP_CORR_AVG, n_cyc)
    n = avg * CODE_SAMPLES
    it = 0
    while freq < MAX_FREQ and it < it_sweep:
        x, _ = wipeoff(data, freq, 0, n, t)
        spec = coherent_spectrum(x, 0, avg)
        hits = []
        for prn in sat_list:
            corr = np.abs(ifft(spec * np.conjugate(spectra[prn])))
            delay, z = peak_test(corr)
            if delay > -1:
                found.append((z, prn, freq, delay))
                hits.append(prn)
        for prn in hits:
            sat_list.remove(prn)
        freq += STEP_FREQ
        if freq >= MAX_FREQ:
            ready = True
            freq -= MAX_FREQ - MIN_FREQ
        it += 1
    return ready, freq, sorted(found, reverse=True)


def fit_code_phase(corr, mx):
    """gpslib.py:1268-1290 — mean of a triangle fit and a parabola fit through
    corr[mx-1], corr[mx], corr[mx+1] (circular neighbours)."""
    n = len(corr)
    lo = mx - 1 if mx > 0 else n - 1
    hi = mx + 1 if mx < n - 1 else 0
    if corr[lo] > corr[hi]:
        tri = 0.5 * (corr[hi] - corr[lo]) / (corr[mx] - corr[hi])
    else:
        tri = 0.5 * (corr[hi] - corr[lo]) / (corr[mx] - corr[lo])
    par = 0.5 * (corr[hi] - corr[lo]) / (2 * corr[mx] - corr[hi] - corr[lo])
    return mx + 0.5 * (tri + par)


class Channel:
    """Signal part of gpslib.SatStream (gpslib.py:1044-1446, 1626-1631): one
    tracked satellite.  Nav-bit/subframe decoding (evalEdges and below) is not
    part of the hot path; `edges` mirrors SatStream.EDGES so that a decoder can
    run on top, and `consume_edges()` applies the bookkeeping logicalBits does
    (gpslib.py:1465-1487)."""

    GAIN_UNLOCKED = 10        # gpslib.py:1046
    GAIN_LOCKED = 1           # gpslib.py:1047
    MIN_CORR_Q = -0.9         # gpslib.py:1048

    def __init__(self, prn, freq, delay=0, it_sweep=IT_SWEEP, corr_min=CORR_MIN,
                 corr_avg=CORR_AVG, sweep_corr_avg=SWEEP_CORR_AVG, n_cyc=32):
        # gpslib.py:1050-1091
        self.prn = prn
        self.n_cyc = n_cyc
        self.ngps = n_cyc * CODE_SAMPLES
        self.t = sec_time(self.ngps)
        self.edges = [0]
        self.locked = False
        self.phase = 0.0
        self.freq = freq
        self.prev_samples = []
        self.ms_time = 0
        self.smp_time = 0
        self.spectrum = code_spectrum(prn)
        self.delay = delay
        self.no_sec = 1024 // n_cyc
        self.corr_min = corr_min
        self.corr_avg = min(corr_avg, n_cyc)
        self.sweep_corr_avg = sweep_corr_avg
        self.code_rep = np.tile(ca_code_2048(prn), n_cyc)       # gpslib.py:81-87, delay 0
        self.std_dev = 0.005
        self.amplitude = 0.0
        self.max_corr = 0.0
        self.sweep = False
        self.it_sweep = it_sweep
        self.prev_stream_no = 0
        self.prev_signal = 0
        self.df = [0]
        self.corr_q = 0
        self.corr_l = 0
        self.corrlst_no = 60 * self.no_sec
        self.corrlst = [0]
        self.rep_sweep = False
        self.rep_sweep_reported = False
        self.prompt = np.zeros(0, dtype=C64)     # last gpsData (CALC_PLOT hook, :1442-1444)
        self.prompt_n0 = []                      # segment start (ST + n0) of every prompt value
        self.new_edges = []                      # edges appended during the last process()
        self.report_due = False
        self.freq_save = None
        self.df_save = None

    # -- state helpers: gpslib.py:1095-1120 ---------------------------------
    def _erase_prev(self):
        self.edges = [0]
        self.prev_samples = []

    def _unlock(self):
        self.locked = False
        self.corrlst = [0]
        self.ms_time = 0
        self.phase = 0.0
        self._erase_prev()

    def _init_sweep(self):
        self._unlock()
        self.freq_save = self.freq
        self.df_save = self.df.copy()
        self.freq = MIN_FREQ
        self.df = [0]
        self.sweep = True

    # -- gpslib.py:1141-1210 ---------------------------------------------------
    def process(self, data, smp_time, sweep=False):
        self.smp_time = smp_time
        self.new_edges = []
        self.report_due = False
        stream_no = smp_time // self.ngps
        if stream_no - 1 != self.prev_stream_no:
            self._erase_prev()
        self.prev_stream_no = stream_no
        sweep = sweep and not self.sweep
        if sweep:
            self._init_sweep()

        if self.sweep:
            self.rep_sweep = self.sweep
            self.sweep, self.freq, self.max_corr, delay, code_phase = self._sweep_frequency(data, self.freq)
            self.corr_q, self.corr_l = self._corr_quality(code_phase)
            if delay >= 0:
                self.delay = delay
            elif not self.sweep:
                self.freq = self.freq_save
                self.df = self.df_save.copy()
            if stream_no % self.no_sec == 0:
                self.report_due = True
                self.rep_sweep_reported = self.rep_sweep
                self.rep_sweep = False
        else:
            data, self.phase = wipeoff(data, self.freq, self.phase, self.ngps, self.t)
            _, delay, code_phase, z = self._code_corr(data, self.corr_avg)
            self.corr_q, self.corr_l = self._corr_quality(code_phase)
            if delay >= 0:
                self.delay = delay
            g = self._decode(data, self.delay)
            self.std_dev = np.std(np.abs(g))
            self.amplitude = np.mean(np.abs(g)) / self.std_dev
            self.max_corr = z
            if stream_no % self.no_sec == 0:
                self.report_due = True
                self.rep_sweep_reported = self.rep_sweep
                self.rep_sweep = False
                if self.locked:
                    self.consume_edges()                     # evalEdges, gpslib.py:1193-1194
                # gpslib.py:1134-1138 (checkCorrQuality)
                if len(self.corrlst) >= self.corrlst_no:
                    sweep = self.corr_q < self.MIN_CORR_Q
            if sweep:
                self._init_sweep()
            else:
                dfreq, shift, self.locked, _ = self._pll(g)
                self.phase += shift
                f = self.freq + dfreq
                self.freq = MAX_FREQ if f > MAX_FREQ else (MIN_FREQ if f < MIN_FREQ else f)  # :1626-1631
        return self.sweep, self.report_due, code_phase, (self.corr_q, self.corr_l)

    # -- gpslib.py:1215-1262 ---------------------------------------------------
    def _pll(self, g):
        max_df = 20 / self.no_sec
        was_locked = self.locked
        n = len(g)
        ph = np.arctan(g.imag / g.real)
        turns = 0
        real_ph = np.copy(ph)
        for i in range(1, n):
            d = ph[i] - ph[i - 1]
            if abs(d) > 2.0:
                turns -= np.sign(d)
            real_ph[i] += turns * np.pi
        offset = np.mean(real_ph[-4:])
        dev = np.mean(real_ph)
        if was_locked:
            df = self.GAIN_LOCKED * dev + np.mean(self.df)
            if abs(df) > max_df:
                df = np.sign(df) * max_df
            if len(self.df) >= self.no_sec:
                del self.df[0]
            self.df.append(df)
        else:
            df = self.GAIN_UNLOCKED * dev
            self.df = [df]
        if abs(dev) < 0.1:
            was_locked = True
        return df, offset, was_locked, real_ph

    # -- gpslib.py:1293-1304, 1315-1327 ----------------------------------------
    def _find(self, corr):
        mean = np.mean(corr)
        std = np.std(corr)
        mx = np.argmax(corr)
        z = (corr[mx] - mean) / std
        if z > self.corr_min:
            return mx, fit_code_phase(corr, mx), z
        return -1, -1.0, z

    def _code_corr(self, data, avg):
        nc = len(data) // CODE_SAMPLES
        p = (nc - avg) // 2
        spec = coherent_spectrum(data, p, avg)
        corr = np.abs(ifft(spec * np.conjugate(self.spectrum)))
        delay, code_phase, z = self._find(corr)
        self.last_corr = corr
        return corr, delay, code_phase, z

    # -- gpslib.py:1331-1339 -----------------------------------------------------
    def _corr_quality(self, code_phase):
        self.corrlst.append(-1 if code_phase < 0 else 1)
        if len(self.corrlst) > self.corrlst_no:
            del self.corrlst[0]
        return np.mean(self.corrlst), np.mean(self.corrlst[-self.no_sec:])

    # -- gpslib.py:1350-1380 -----------------------------------------------------
    def _corr_max(self, data, avg, freq):
        x, _ = wipeoff(data, freq, 0, avg * CODE_SAMPLES, self.t)
        corr, delay, code_phase, z = self._code_corr(x, avg)
        return delay, (corr[delay] if delay > -1 else 0), code_phase, z

    def _sweep_frequency(self, data, freq):
        running = True
        j = 0
        delay = -1
        code_phase = -1
        while delay < 0 and j < self.it_sweep:
            delay, _, code_phase, z = self._corr_max(data, self.sweep_corr_avg, freq)
            if delay < 0:
                freq = freq + STEP_FREQ
            j += 1
        if delay >= 0:
            running = False
        elif freq > MAX_FREQ:
            freq = MIN_FREQ
            running = False
        return running, freq, z, delay, code_phase

    # -- gpslib.py:1394-1446 (decodeData) ------------------------------------------
    def _decode(self, data, delay):
        min_edge = 3 * self.std_dev
        prev_sign = (2 * (len(self.edges) % 2) - 1) * self.edges[0]
        y = np.roll(self.code_rep, delay) * data
        nps = len(self.prev_samples)
        if nps > 0:
            y = np.append(self.prev_samples, y)
        ns = self.ngps + nps
        n0 = 0
        n1 = nps + delay
        if n1 == 0:
            n1 = CODE_SAMPLES
            st = self.smp_time
        else:
            st = self.smp_time + delay - CODE_SAMPLES
        out = []
        self.prompt_n0 = []
        while n1 <= ns:
            m = np.mean(y[n0:n1])
            out.append(m)
            self.prompt_n0.append(st + n0)
            if self.locked:
                s = np.sign(m.real)
                if self.edges[0] == 0:
                    self.edges[0] = s
                    prev_sign = s
                elif (s != prev_sign and prev_sign * self.prev_signal > 0
                      and abs(m.real - self.prev_signal) > min_edge):
                    self.edges.append((self.ms_time, st + n0))
                    self.new_edges.append((self.ms_time, st + n0))
                    prev_sign = s
                self.prev_signal = m.real
                self.ms_time += 1
            n0 = n1
            n1 += CODE_SAMPLES
        out = np.asarray(out, dtype=C64)
        self.prev_samples = y[n0:ns]
        self.prompt = out
        return out

    def consume_edges(self):
        """Bookkeeping that evalEdges -> logicalBits applies to EDGES once per
        second while locked (gpslib.py:1455-1487): keep [last sign, last edge]."""
        n = len(self.edges)
        if n > 2:
            last = self.edges[0] * (-1) ** (n - 2)
            self.edges = [last, self.edges[-1]]


# ---------------------------------------------------------------------------
# Generalised acquisition grid (reference-derived; BASELINE.json configs 2 & 4)
# ---------------------------------------------------------------------------
ACQ_MODE_ABS = 0      # statistic on |c| (the reference's, gpsrecv.py:258), nnoncoh must be 1
ACQ_MODE_POW = 1      # statistic on sum_k |c_k|^2 (non-coherent accumulation)
SECOND_PEAK_GUARD = 4  # samples (= 2 chips) excluded around the peak for the 2nd-peak search


def acq_cell_stats(stat):
    """argmax / mean / population std / z / neighbours / second peak of one
    2048-lag statistic (gpsrecv.py:217-227 plus the peak-to-second-peak ratio
    named in BASELINE.json)."""
    n = len(stat)
    mx = int(np.argmax(stat))
    mean = float(np.mean(stat))
    std = float(np.std(stat))
    peak = float(stat[mx])
    lag = np.arange(n)
    dist = np.minimum((lag - mx) % n, (mx - lag) % n)
    second = float(np.max(stat[dist > SECOND_PEAK_GUARD]))
    return dict(mx=mx, peak=peak, mean=mean, std=std, z=(peak - mean) / std,
                em1=float(stat[(mx - 1) % n]), ep1=float(stat[(mx + 1) % n]), second=second)


def acq_grid(data, prns, f0, fstep, nbins, tcoh_ms, nnoncoh, mode=ACQ_MODE_POW, spectra=None):
    """Acquisition statistic over PRN x Doppler x code phase.

    Follows gpsrecv.py:232-235 (wipe-off with t = (n+1)/fs as float32, phase 0,
    continuous over the whole tcoh*nnoncoh span), :250-258 (mean of tcoh 1-ms
    FFTs, times conj(code spectrum), ifft, abs) and :217-227 (argmax, mean,
    std), adding acc += |c|^2 over the nnoncoh intervals (ACQ_MODE_POW).
    Returns a dict of arrays shaped [len(prns), nbins]."""
    n_ms = tcoh_ms * nnoncoh
    n = n_ms * CODE_SAMPLES
    t = sec_time(n)
    keys = ("mx", "peak", "mean", "std", "z", "em1", "ep1", "second")
    out = {k: np.zeros((len(prns), nbins), dtype=np.int32 if k == "mx" else np.float64) for k in keys}
    if spectra is None:
        spectra = {p: code_spectrum(p) for p in prns}
    for b in range(nbins):
        freq = f0 + b * fstep
        x, _ = wipeoff(data, freq, 0, n, t)
        specs = [coherent_spectrum(x, k * tcoh_ms, tcoh_ms) for k in range(nnoncoh)]
        for i, p in enumerate(prns):
            cs = np.conjugate(spectra[p])
            if mode == ACQ_MODE_ABS:
                stat = np.abs(ifft(specs[0] * cs))
            else:
                stat = np.zeros(CODE_SAMPLES)
                for s in specs:
                    c = ifft(s * cs)
                    stat += c.real * c.real + c.imag * c.imag
            st = acq_cell_stats(stat)
            for k in keys:
                out[k][i, b] = st[k]
    return out
