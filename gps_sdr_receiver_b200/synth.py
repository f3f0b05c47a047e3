"""Synthetic RTL-SDR style uint8 I/Q recordings with known GPS L1 C/A satellites.

This is measurement/test infrastructure, not part of the
product path.  The byte layout is the one the reference reader expects
(/root/reference/src/gpsrecv.py:168-173): one sample = two bytes, byte 0 = I,
byte 1 = Q, value = round((x + 1) * 127.5) clipped to [0, 255].

Signal model per satellite (sample index n counted from the start of the
recording, fs = 2.048 MS/s, 2048 samples per 1 ms code period):

    a * chip[floor(((n - tau) mod 2048) * 1023 / 2048)] * navbit(n)
      * exp(j * (2 pi * (f n' + 0.5 * fdot * n'^2) + phi0)),   n' = (n + 1) / fs

The Gold codes come from the G1/G2 shift registers (IS-GPS-200), not from the
reference's literal table; tests check that both agree.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FS = 2_048_000
CODE_SAMPLES = 2048

# G2 output taps (1-based register stages) for PRN 1..37, IS-GPS-200 table 3-Ia.
G2_TAPS = {
    1: (2, 6), 2: (3, 7), 3: (4, 8), 4: (5, 9), 5: (1, 9), 6: (2, 10), 7: (1, 8),
    8: (2, 9), 9: (3, 10), 10: (2, 3), 11: (3, 4), 12: (5, 6), 13: (6, 7),
    14: (7, 8), 15: (8, 9), 16: (9, 10), 17: (1, 4), 18: (2, 5), 19: (3, 6),
    20: (4, 7), 21: (5, 8), 22: (6, 9), 23: (1, 3), 24: (4, 6), 25: (5, 7),
    26: (6, 8), 27: (7, 9), 28: (8, 10), 29: (1, 6), 30: (2, 7), 31: (3, 8),
    32: (4, 9), 33: (5, 10), 34: (4, 10), 35: (1, 7), 36: (2, 8), 37: (4, 10),
}


def gold_chips(prn: int) -> np.ndarray:
    """1023 chips of the C/A code of `prn` as int8, logical 1 -> +1, 0 -> -1."""
    t1, t2 = G2_TAPS[prn]
    g1 = [1] * 10
    g2 = [1] * 10
    out = np.empty(1023, dtype=np.int8)
    for i in range(1023):
        bit = g1[9] ^ g2[t1 - 1] ^ g2[t2 - 1]
        out[i] = 1 if bit else -1
        f1 = g1[2] ^ g1[9]
        f2 = g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9]
        g1 = [f1] + g1[:9]
        g2 = [f2] + g2[:9]
    return out


@dataclass
class Sat:
    prn: int
    doppler: float            # Hz at n = 0
    delay: float              # code phase in samples, 0 <= delay < 2048
    amp: float = 0.07
    phi0: float = 0.3         # rad
    doppler_rate: float = 0.0  # Hz/s
    bit_offset_ms: int = 7    # first nav-bit boundary (ms since start)
    bit_seed: int = 0         # seed of the nav-bit stream
    bits: np.ndarray | None = field(default=None, repr=False)  # optional +-1 bits, indexed by the ABSOLUTE bit number
                                                               # (modulo their length): e.g. 2 * navbits.encode_frames(..) - 1


def _nav_bits(sat: Sat, nbits: int) -> np.ndarray:
    if sat.bits is not None:
        reps = -(-nbits // len(sat.bits))
        return np.tile(np.asarray(sat.bits, dtype=np.int8), reps)[:nbits]
    rng = np.random.default_rng(1000 + 7919 * sat.prn + sat.bit_seed)
    return (2 * rng.integers(0, 2, nbits) - 1).astype(np.int8)


def make_iq(sats: list[Sat], n_ms: int, noise_sigma: float = 0.25, seed: int = 1,
            start_sample: int = 0, as_float: bool = False) -> np.ndarray:
    """Return uint8[2 * n_ms * 2048] interleaved I,Q (or the float complex128
    pre-quantisation signal when `as_float`)."""
    n = n_ms * CODE_SAMPLES
    idx = np.arange(start_sample, start_sample + n, dtype=np.int64)
    tt = (idx + 1).astype(np.float64) / FS
    x = np.zeros(n, dtype=np.complex128)
    for s in sats:
        chips = gold_chips(s.prn).astype(np.float64)
        # the code phase is measured in receiver samples: the code start sits
        # at n = delay (mod 2048); chips advance at 1023/2048 chip per sample
        ph = np.mod(idx.astype(np.float64) - s.delay, CODE_SAMPLES)
        ci = np.floor(ph * (1023.0 / CODE_SAMPLES)).astype(np.int64) % 1023
        code = chips[ci]
        # nav bits: 20 ms per bit, boundaries aligned with code starts
        code_no = np.floor((idx.astype(np.float64) - s.delay) / CODE_SAMPLES).astype(np.int64)
        bit_no = np.floor_divide(code_no - s.bit_offset_ms, 20)
        if s.bits is not None:               # explicit message: consistent when a recording is generated in pieces
            nav = np.asarray(s.bits, dtype=np.float64)[np.mod(bit_no, len(s.bits))]
        else:
            nb = int(bit_no.max() - bit_no.min()) + 1
            bits = _nav_bits(s, nb + 4)
            nav = bits[bit_no - bit_no.min()].astype(np.float64)
        phase = 2.0 * np.pi * (s.doppler * tt + 0.5 * s.doppler_rate * tt * tt) + s.phi0
        x += s.amp * code * nav * np.exp(1j * phase)
    rng = np.random.default_rng(seed + 31 * (start_sample // CODE_SAMPLES))
    x += noise_sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    if as_float:
        return x
    q = np.empty(2 * n, dtype=np.float64)
    q[0::2] = x.real
    q[1::2] = x.imag
    q = np.clip(np.rint((q + 1.0) * 127.5), 0, 255)
    return q.astype(np.uint8)


def default_constellation(nsat: int = 6, seed: int = 5) -> list[Sat]:
    """A fixed, well separated set of satellites (PRN >= 2: the reference's
    cold-start search skips PRN 1, gpsrecv.py:36)."""
    rng = np.random.default_rng(seed)
    prns = rng.permutation(np.arange(2, 33))[:nsat]
    sats = []
    for k, p in enumerate(sorted(int(v) for v in prns)):
        sats.append(Sat(prn=p,
                        doppler=float(np.round(rng.uniform(-4200, 4200), 1)),
                        delay=float(np.round(rng.uniform(3, 2040), 2)),
                        amp=float(np.round(rng.uniform(0.06, 0.085), 3)),
                        phi0=float(np.round(rng.uniform(-3, 3), 2)),
                        doppler_rate=float(np.round(rng.uniform(-1.5, 1.5), 2)),
                        bit_offset_ms=int(rng.integers(0, 20)),
                        bit_seed=k))
    return sats


def make_iq_dev(sats: list[Sat], n_ms: int, noise_sigma: float = 0.25, seed: int = 1, start_sample: int = 0,
                out=None, device: int = 0, stream=None):
    """Same signal model generated on the GPU straight into HBM (csrc/gr_synth.cu; different
    random streams than make_iq).  Returns a torch uint8 tensor [2 * n_ms * 2048]."""
    import ctypes as C
    import torch
    from . import _capi
    _capi.init(device)
    n = n_ms * CODE_SAMPLES
    if out is None:
        out = torch.empty(2 * n, dtype=torch.uint8, device=f"cuda:{device}")
    arr = (_capi.SynthSat * max(1, len(sats)))()
    for i, s in enumerate(sats):
        arr[i] = _capi.SynthSat(s.prn, s.bit_offset_ms, s.bit_seed, s.amp, s.doppler, s.doppler_rate, s.delay, s.phi0)
    st = torch.cuda.current_stream(out.device).cuda_stream if stream is None else stream
    _capi.check(_capi.lib().gr_synth_iq_dev(out.data_ptr(), n, int(start_sample), C.addressof(arr), len(sats),
                                            float(noise_sigma), int(seed), st))
    return out
