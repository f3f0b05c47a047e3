"""Geometry-consistent synthetic GPS constellation: orbits -> flight time versus
receiver time -> code delay, carrier phase and nav bits of every visible satellite.  Measurement / test
infrastructure: it feeds the device generator `gr_synth_geo_dev` (csrc/gr_synth.cu)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import navbits
from .position import C_LIGHT, F_L1, elevation, flight_time, geo_to_ecef, sat_clock, sat_ecef

NODE_DT = 0.1          # s between flight-time nodes (cubic interpolation on the device)


@dataclass
class GeoSat:
    prn: int
    eph: dict
    amp: float
    tau: np.ndarray = field(repr=False)       # t_rx - t_sat_clock at the nodes, s (includes the satellite clock error)
    bits: np.ndarray = field(repr=False)      # int8 0/1, bit k covers satellite-clock time [bit_t0 + 20 ms k, ...)
    bit_t0: float = 0.0


def make_ephemeris(prn: int, toe: int, omega0: float, m0: float, rng) -> dict:
    """Plausible GPS orbit (a = 26 560 km, i = 55 deg, small eccentricity) quantised to the LNAV LSBs."""
    raw = dict(weekNum=300 + prn % 7, satAcc=1, satHealth=0, Tgd=float(rng.integers(-20, 20)) * 2.0 ** -31,
               IODC=prn, Toc=toe, af2=0.0, af1=float(rng.integers(-40, 40)) * 2.0 ** -43,
               af0=float(rng.integers(-200000, 200000)) * 2.0 ** -31,
               IODE2=prn, Crs=float(rng.uniform(-100, 100)), deltaN=float(rng.uniform(3e-9, 6e-9)), M0=m0,
               Cuc=float(rng.uniform(-5e-6, 5e-6)), e=float(rng.uniform(0.001, 0.015)), Cus=float(rng.uniform(-9e-6, 9e-6)),
               sqrtA=5153.6 + float(rng.uniform(-0.3, 0.3)), Toe=toe,
               Cic=float(rng.uniform(-2e-7, 2e-7)), omegaBig=omega0, Cis=float(rng.uniform(-2e-7, 2e-7)),
               i0=math.radians(55.0) + float(rng.uniform(-0.02, 0.02)), IODE3=prn, Crc=float(rng.uniform(150, 350)),
               omegaSmall=float(rng.uniform(-3, 3)), omegaDot=float(rng.uniform(-8.6e-9, -7.6e-9)),
               IDOT=float(rng.uniform(-5e-10, 5e-10)))
    # what the receiver will decode is the quantised message: make the truth exactly that
    eph = {}
    for k in (1, 2, 3):
        eph.update(navbits.decode_subframe(navbits.encode_subframe(k, 1, raw))[1])
    for k in ("ID", "tow"):
        eph.pop(k)
    return eph


def build(rx_geo=(49.0830, 8.3076, 120.0), tow0: int = 345600, seconds: float = 24.0, n_sat: int = 7, amp: float = 0.07,
          min_elev: float = 15.0, seed: int = 1, rx_clock_bias: float = 1.2345e-4):
    """Satellites above `min_elev` for a static receiver; the recording starts at GPS time of week `tow0`
    (receiver clock = GPS time + rx_clock_bias).  Returns (rx_ecef, [GeoSat])."""
    rng = np.random.default_rng(seed)
    rx = geo_to_ecef(*rx_geo)
    toe = (tow0 // 7200) * 7200
    n_nodes = int(math.ceil(seconds / NODE_DT)) + 4
    sats, prn = [], 1
    tries = 0
    while len(sats) < n_sat and tries < 4000:
        tries += 1
        eph = make_ephemeris(prn, toe, float(rng.uniform(-math.pi, math.pi)), float(rng.uniform(-math.pi, math.pi)), rng)
        p, _ = sat_ecef(eph, tow0)
        if elevation(rx, p) < min_elev:
            continue
        tau = np.empty(n_nodes)
        for i in range(n_nodes):
            t_rx = tow0 + (i - 1) * NODE_DT                          # node 0 sits one step before the start
            fl = flight_time(eph, rx, t_rx)
            t_sys = t_rx - fl
            _, rel = sat_ecef(eph, t_sys)
            # the satellite's own clock reads t_sys + clock error; the receiver clock reads t_rx + bias
            tau[i] = (t_rx + rx_clock_bias) - (t_sys + sat_clock(eph, t_sys, rel))
        # nav message by satellite-clock time: subframe boundaries at multiples of 6 s
        first_sub = int(math.floor((tow0 - 1.0) / 6.0))              # subframe containing the start (with margin)
        n_sub = int(math.ceil((seconds + 2.0) / 6.0)) + 1
        bits = np.concatenate([np.asarray(navbits.encode_subframe((first_sub + k) % 5 + 1, first_sub + k + 1, eph), dtype=np.int8)
                               for k in range(n_sub)])
        sats.append(GeoSat(prn=prn, eph=eph, amp=amp, tau=tau, bits=bits, bit_t0=first_sub * 6.0))
        prn += 1
    if len(sats) < n_sat:
        raise RuntimeError("could not place enough visible satellites")
    return rx, sats


def doppler_at_start(s: GeoSat) -> float:
    """Carrier Doppler (Hz) at the first sample: -f_L1 d(tau)/dt."""
    return -F_L1 * (s.tau[2] - s.tau[0]) / (2 * NODE_DT)


def code_delay_at_start(s: GeoSat, tow0: float, rx_clock_bias: float) -> float:
    """Receiver sample offset (mod 2048) of the first code start at / after sample 0."""
    # receiver clock reading at sample n: tow0 + bias + n / fs; satellite clock = that - tau
    t_sat0 = tow0 + rx_clock_bias - s.tau[1]
    frac_ms = (t_sat0 * 1e3) % 1.0
    return ((1.0 - frac_ms) % 1.0) * 2048.0


def make_iq_dev(sats: list[GeoSat], n_ms: int, tow0: float, rx_clock_bias: float, noise_sigma: float = 0.25, seed: int = 1,
                device: int = 0, out=None):
    """The recording of `build`'s constellation, generated on the GPU (csrc/gr_synth.cu, gr_synth_geo_dev):
    torch uint8 [2 * n_ms * 2048], sample 0 at receiver clock reading tow0 + rx_clock_bias."""
    import ctypes as C
    import torch
    from . import _capi
    _capi.init(device)
    n = n_ms * 2048
    if out is None:
        out = torch.empty(2 * n, dtype=torch.uint8, device=f"cuda:{device}")
    keep = []
    arr = (_capi.SynthGeoSat * len(sats))()
    for i, s in enumerate(sats):
        d_tau = torch.from_numpy(np.ascontiguousarray(s.tau, dtype=np.float64)).to(out.device)
        d_bits = torch.from_numpy(np.ascontiguousarray(s.bits, dtype=np.int8)).to(out.device)
        keep += [d_tau, d_bits]
        arr[i] = _capi.SynthGeoSat(s.prn, len(s.tau), len(s.bits), s.amp, int(round(s.bit_t0 * 1000)), d_tau.data_ptr(), d_bits.data_ptr())
    t0 = tow0 + rx_clock_bias
    t0_ms = int(math.floor(t0 * 1000.0))
    t0_frac = t0 - t0_ms * 1e-3
    st = torch.cuda.current_stream(out.device).cuda_stream
    _capi.check(_capi.lib().gr_synth_geo_dev(out.data_ptr(), n, 0, C.addressof(arr), len(sats), t0_ms, t0_frac, NODE_DT,
                                             float(noise_sigma), int(seed), st))
    torch.cuda.synchronize()
    return out


def make_iq_host(sats: list[GeoSat], n_ms: int, tow0: float, rx_clock_bias: float, noise_sigma: float = 0.25, seed: int = 1,
                 piece_ms: int = 250) -> np.ndarray:
    """numpy twin of make_iq_dev (same signal model, float64 math, numpy noise): slow (about 5 s of CPU per second
    of recording and satellite); for CPU-only experiments such as oracle/e2e_reference_fix.py."""
    from .synth import gold_chips
    out = np.empty(2 * n_ms * 2048, dtype=np.uint8)
    t0 = tow0 + rx_clock_bias
    t0_ms = int(math.floor(t0 * 1000.0))
    t0_frac = t0 - t0_ms * 1e-3
    chips = {s.prn: gold_chips(s.prn).astype(np.float64) for s in sats}
    rng = np.random.default_rng(seed)
    for m0 in range(0, n_ms, piece_ms):
        m1 = min(n_ms, m0 + piece_ms)
        n = np.arange(m0 * 2048, m1 * 2048, dtype=np.int64)
        trel = n.astype(np.float64) / 2048000.0
        x = trel / NODE_DT + 1.0
        sig = np.zeros(n.size, dtype=np.complex128)
        for s in sats:
            k = np.clip(np.floor(x).astype(np.int64), 1, len(s.tau) - 3)
            u = x - k
            y0, y1, y2, y3 = s.tau[k - 1], s.tau[k], s.tau[k + 1], s.tau[k + 2]
            tau = (-u * (u - 1) * (u - 2) / 6 * y0 + (u + 1) * (u - 1) * (u - 2) / 2 * y1
                   - (u + 1) * u * (u - 2) / 2 * y2 + (u + 1) * u * (u - 1) / 6 * y3)
            ysat = (t0_frac + (trel - tau)) * 1e3
            ms_f = np.floor(ysat)
            ci = np.minimum((ysat - ms_f) * 1023.0, 1022.999).astype(np.int64)
            b = np.floor_divide(t0_ms + ms_f.astype(np.int64) - int(round(s.bit_t0 * 1000)), 20)
            nav = 2.0 * s.bits[np.clip(b, 0, len(s.bits) - 1)].astype(np.float64) - 1.0
            cyc = -F_L1 * tau
            sig += s.amp * chips[s.prn][ci] * nav * np.exp(2j * np.pi * (cyc - np.floor(cyc)))
        sig += noise_sigma * (rng.standard_normal(n.size) + 1j * rng.standard_normal(n.size))
        q = np.empty(2 * n.size)
        q[0::2], q[1::2] = sig.real, sig.imag
        out[2 * m0 * 2048:2 * m1 * 2048] = np.clip(np.rint((q + 1.0) * 127.5), 0, 255).astype(np.uint8)
    return out
