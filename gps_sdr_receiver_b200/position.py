"""From tracked observables to a position fix -- host float64, a few calls per
second, deliberately simple: it exists so that the whole chain (synthetic constellation -> acquisition
-> tracking -> nav bits -> pseudoranges -> fix) can be closed and checked inside this repository and
on the GPU box, where the reference's own gpseval / SatOrbit / leastSquaresPos code is not available.
In a deployment the reference's consumer runs unchanged on the same frame dicts and code phases.

Conventions: IS-GPS-200 (user algorithm of table 20-IV for the satellite position, the relativistic
clock term, Sagnac rotation during the signal's flight); field names are those of the frame dicts
(`navbits.decode_subframe`, i.e. `gpslib.Subframe`, src/gpslib.py:316-371).
"""
from __future__ import annotations

import math

import numpy as np

from . import glob

C_LIGHT = 2.99792458e8
MU = 3.986005e14                 # WGS-84 value used by GPS, m^3/s^2
OMEGA_E = 7.2921151467e-5        # rad/s
F_REL = -4.442807633e-10         # s / sqrt(m)
F_L1 = 1575.42e6
FS = 2_048_000.0
WGS_A, WGS_F = 6378137.0, 1.0 / 298.257223563


def kepler_E(M: float, e: float) -> float:
    E = M
    for _ in range(12):
        E = E - (E - e * math.sin(E) - M) / (1.0 - e * math.cos(E))
    return E


def sat_ecef(eph: dict, t: float) -> tuple[np.ndarray, float]:
    """ECEF position of the satellite at GPS time of week `t` (s) and its relativistic clock term (s)."""
    A = eph["sqrtA"] ** 2
    n = math.sqrt(MU / A ** 3) + eph["deltaN"]
    tk = t - eph["Toe"]
    tk = tk - 604800.0 if tk > 302400.0 else (tk + 604800.0 if tk < -302400.0 else tk)
    E = kepler_E(eph["M0"] + n * tk, eph["e"])
    nu = math.atan2(math.sqrt(1.0 - eph["e"] ** 2) * math.sin(E), math.cos(E) - eph["e"])
    phi = nu + eph["omegaSmall"]
    s2, c2 = math.sin(2 * phi), math.cos(2 * phi)
    u = phi + eph["Cus"] * s2 + eph["Cuc"] * c2
    r = A * (1.0 - eph["e"] * math.cos(E)) + eph["Crs"] * s2 + eph["Crc"] * c2
    inc = eph["i0"] + eph["IDOT"] * tk + eph["Cis"] * s2 + eph["Cic"] * c2
    om = eph["omegaBig"] + (eph["omegaDot"] - OMEGA_E) * tk - OMEGA_E * eph["Toe"]
    xp, yp = r * math.cos(u), r * math.sin(u)
    pos = np.array([xp * math.cos(om) - yp * math.cos(inc) * math.sin(om),
                    xp * math.sin(om) + yp * math.cos(inc) * math.cos(om),
                    yp * math.sin(inc)])
    return pos, F_REL * eph["e"] * eph["sqrtA"] * math.sin(E)


def sat_clock(eph: dict, t: float, rel: float) -> float:
    dt = t - eph.get("Toc", 0)
    return eph.get("af0", 0.0) + eph.get("af1", 0.0) * dt + eph.get("af2", 0.0) * dt * dt + rel - eph.get("Tgd", 0.0)


def geo_to_ecef(lat_deg: float, lon_deg: float, h: float) -> np.ndarray:
    lat, lon = math.radians(lat_deg), math.radians(lon_deg)
    e2 = WGS_F * (2.0 - WGS_F)
    N = WGS_A / math.sqrt(1.0 - e2 * math.sin(lat) ** 2)
    return np.array([(N + h) * math.cos(lat) * math.cos(lon), (N + h) * math.cos(lat) * math.sin(lon),
                     (N * (1.0 - e2) + h) * math.sin(lat)])


def ecef_to_geo(p) -> tuple[float, float, float]:
    x, y, z = (float(v) for v in p)
    e2 = WGS_F * (2.0 - WGS_F)
    lon = math.atan2(y, x)
    rho = math.hypot(x, y)
    lat = math.atan2(z, rho * (1.0 - e2))
    for _ in range(8):
        N = WGS_A / math.sqrt(1.0 - e2 * math.sin(lat) ** 2)
        h = rho / math.cos(lat) - N
        lat = math.atan2(z, rho * (1.0 - e2 * N / (N + h)))
    return math.degrees(lat), math.degrees(lon), h


def elevation(rx: np.ndarray, sat: np.ndarray) -> float:
    up = rx / np.linalg.norm(rx)
    los = sat - rx
    return math.degrees(math.asin(float(los @ up) / float(np.linalg.norm(los))))


def flight_time(eph: dict, rx: np.ndarray, t_rx: float) -> float:
    """Signal flight time to a receiver at ECEF `rx` (fixed to the Earth) for reception at GPS time t_rx."""
    tau = 0.075
    for _ in range(6):
        ps, _ = sat_ecef(eph, t_rx - tau)
        a = OMEGA_E * tau                                   # the Earth turned by a during the flight
        ps_r = np.array([math.cos(a) * ps[0] + math.sin(a) * ps[1], -math.sin(a) * ps[0] + math.cos(a) * ps[1], ps[2]])
        tau = float(np.linalg.norm(ps_r - rx)) / C_LIGHT
    return tau


def solve_fix(ephs: list[dict], t_tx: np.ndarray, t_rx_local: float, x0=None, iters: int = 10):
    """Gauss-Newton for (x, y, z, receiver clock bias) from the satellites' transmit times (their own
    clocks, GPS time of week) of the signal received at local receiver time `t_rx_local` (s, arbitrary
    offset).  Returns (ecef[3], clock_bias_s, residuals_m)."""
    n = len(ephs)
    if n < 4:
        raise ValueError("need at least 4 satellites")
    sat_pos, t_corr = [], []
    for eph, tt in zip(ephs, t_tx):
        _, rel = sat_ecef(eph, tt)
        tsys = tt - sat_clock(eph, tt, rel)                    # satellite clock reading -> GPS system time
        p, _ = sat_ecef(eph, tsys)
        sat_pos.append(p)
        t_corr.append(tsys)
    sat_pos, t_corr = np.array(sat_pos), np.array(t_corr)
    pr = (t_rx_local - t_corr) * C_LIGHT                       # pseudoranges with the common receiver offset
    x = np.zeros(4) if x0 is None else np.array(list(x0) + [0.0])[:4]
    x[3] = float(np.mean(pr)) - 2.2e7
    for _ in range(iters):
        H, r = np.zeros((n, 4)), np.zeros(n)
        for i in range(n):
            tau = (pr[i] - x[3]) / C_LIGHT
            a = OMEGA_E * tau
            p = sat_pos[i]
            pr_i = np.array([math.cos(a) * p[0] + math.sin(a) * p[1], -math.sin(a) * p[0] + math.cos(a) * p[1], p[2]])
            d = pr_i - x[:3]
            rho = float(np.linalg.norm(d))
            H[i, :3], H[i, 3] = -d / rho, 1.0
            r[i] = pr[i] - (rho + x[3])
        dx = np.linalg.lstsq(H, r, rcond=None)[0]
        x += dx
        if float(np.linalg.norm(dx[:3])) < 1e-4:
            break
    return x[:3], x[3] / C_LIGHT, r


class ChannelObservables:
    """Transmit-time bookkeeping of one tracked channel: subframe starts (TOW, sample time of the preamble)
    from the frame dicts plus the per-epoch code phase give the satellite's clock reading for the code start
    nearest to any receiver sample time."""

    def __init__(self, prn: int, n_cyc: int):
        self.prn, self.ngps = prn, n_cyc * 2048
        self.eph: dict = {}
        self.ref: tuple[int, float] | None = None      # (sample time of a subframe start, its transmit time of week)
        self.code_phase: list[tuple[int, float, float]] = []   # (epoch sample time, codePhase, FREQ)

    def add_frames(self, frames: list[dict]):
        for f in frames:
            if "ID" not in f:
                continue
            self.eph.update({k: v for k, v in f.items() if k not in ("ID", "tow", "ST", "SAT", "AMP", "CRM", "FRQ", "SWP")})
            self.eph[f"have{f['ID']}"] = True
            # the HOW carries the TOW count of the NEXT subframe: this one started at (tow - 1) * 6 s
            self.ref = (int(f["ST"]), (int(f["tow"]) - 1) * 6.0)

    def add_epoch(self, smp_time: int, code_phase: float, freq: float):
        if code_phase >= 0:
            self.code_phase.append((int(smp_time), float(code_phase), float(freq)))

    @property
    def ready(self) -> bool:
        return self.ref is not None and all(self.eph.get(f"have{i}") for i in (1, 2, 3)) and len(self.code_phase) > 0

    def transmit_time(self, s_rx: float, n_avg: int = 8, centre_ms: float | None = None) -> float:
        """Satellite clock reading (time of week) of the signal that arrives at receiver sample time s_rx
        (SMP_TIME convention).  `centre_ms`: middle of the correlation window inside an epoch
        ((n_cyc - corr_avg) / 2 + corr_avg / 2 blocks, gpslib.py:1315-1327): the time the code phase refers to.
        Default: derived from this channel's N_CYC.  The code phases of the last `n_avg` epochs up to s_rx are each propagated to s_rx and averaged."""
        if centre_ms is None:                                    # 16 ms at N_CYC = 32, 8 at 16, 4 at 8 (CORR_AVG = 8)
            n_cyc = self.ngps // 2048
            corr_avg = min(glob.CORR_AVG, n_cyc)
            centre_ms = (n_cyc - corr_avg) // 2 + corr_avg / 2.0
        st_ref, t_ref = self.ref
        pts = [c for c in self.code_phase if c[0] + centre_ms * 2048.0 <= s_rx][-n_avg:]
        if not pts:
            pts = self.code_phase[:1]
        est = []
        for smp, cp, freq in pts:
            rate = 1.0 + freq / F_L1                             # satellite seconds per receiver second
            s_cs = smp + centre_ms * 2048.0 + cp * (2.0 - rate)  # receive time of the code start of the centre block
            n_ms = round((s_cs - st_ref) / 2048.0)
            t_cs = t_ref + 1e-3 * n_ms                           # its transmit time (satellite clock)
            est.append(t_cs + (s_rx - s_cs) / FS * rate)
        return float(np.mean(est))
