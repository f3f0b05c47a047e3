"""Host-side geometry of the synthetic constellation and the position solver (no GPU): exact transmit
times at the true receiver position must solve back to it."""
from __future__ import annotations

import numpy as np

from gps_sdr_receiver_b200 import constellation as con, position as pos


def test_geodetic_round_trip():
    for lat, lon, h in [(49.083, 8.3076, 120.0), (-33.9, 151.2, 5.0), (0.0, -179.9, 4000.0), (89.0, 10.0, -50.0)]:
        la, lo, hh = pos.ecef_to_geo(pos.geo_to_ecef(lat, lon, h))
        assert abs(la - lat) < 1e-9 and abs(lo - lon) < 1e-9 and abs(hh - h) < 1e-4


def test_constellation_is_visible_and_solves_back_to_the_receiver():
    tow0, bias = 345597, 1.2345e-4
    rx, sats = con.build(seconds=12.0, n_sat=6, tow0=tow0, rx_clock_bias=bias, seed=1)
    assert len({s.prn for s in sats}) == 6
    for s in sats:
        assert pos.elevation(rx, pos.sat_ecef(s.eph, tow0)[0]) >= 15.0
        assert abs(con.doppler_at_start(s)) < 5000.0 and 0.06 < s.tau[1] - bias + pos.sat_clock(s.eph, tow0, 0.0) < 0.09
        # the message the receiver will decode is exactly the ephemeris the geometry was computed from
        from gps_sdr_receiver_b200 import navbits
        for k in range(len(s.bits) // 300):
            st, f = navbits.decode_subframe(s.bits[300 * k:300 * (k + 1)])
            assert st == 0 and all(s.eph[name] == v for name, v in f.items() if name in s.eph)
    for node in (1, 41, 101):                                    # 0 s, 4 s, 10 s into the recording
        t_rx_clock = tow0 + (node - 1) * con.NODE_DT + bias
        ttx = np.array([t_rx_clock - s.tau[node] for s in sats])
        p, cb, res = pos.solve_fix([s.eph for s in sats], ttx, t_rx_clock)
        assert np.linalg.norm(p - rx) < 0.05 and abs(cb - bias) < 1e-9 and np.abs(res).max() < 0.05
