"""BASELINE configs[0] stand-in, reference side (build container only; RUNS THE REAL REFERENCE on the CPU).

A geometry-consistent synthetic recording (gps_sdr_receiver_b200.constellation, numpy twin of the device
generator) is tracked by the unmodified `gpslib.SatStream` and evaluated by the unmodified
`gpseval.prepCodePhase / evalData / ecefPositions` (SatOrbit, leastSquaresPos ...), exactly as
gpsrecv.processData / gpseval.processData chain them (src/gpsrecv.py:492-519, src/gpseval.py:529-540).
On the SAME SatStream outputs this repository's consumer (navbits.FrameDecoder fed with the same EDGES,
position.ChannelObservables / solve_fix) computes its fix.  Printed and stored in
tests/golden/e2e_reference_fix.json: the true position, the reference's fixes, our fix, their distance.

    python oracle/e2e_reference_fix.py
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference/src")

N_CYC, SECONDS, TOW0, BIAS = 32, 26, 345597, 1.2345e-4


def import_gpseval():
    """gpseval imports matplotlib / gpsui / gpxpy at module level (GUI only): stub them."""
    for name in ("matplotlib", "matplotlib.pyplot", "gpsui", "gpxpy"):
        m = types.ModuleType(name)
        m.use = lambda *a, **k: None
        m.ion = lambda *a, **k: None
        sys.modules.setdefault(name, m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import gpseval
    return gpseval


def main():
    from gps_sdr_receiver_b200 import constellation as con, navbits, position as pos
    from gps_sdr_receiver_b200.tracking import new_edges                         # noqa: F401 (documentation of the EDGES format)
    import gpslib                                                               # the reference
    gpseval = import_gpseval()
    assert gpslib.N_CYC == N_CYC if hasattr(gpslib, "N_CYC") else True

    rx, sats = con.build(seconds=SECONDS, n_sat=6, tow0=TOW0, rx_clock_bias=BIAS, seed=1)
    n_ms = SECONDS * 1000 // N_CYC * N_CYC
    raw = con.make_iq_host(sats, n_ms, TOW0, BIAS, noise_sigma=0.25, seed=11)
    ngps = N_CYC * 2048
    n_ep = n_ms // N_CYC

    class Tap(gpslib.SatStream):                      # the unmodified class; only records what evalEdges was given
        tap = None

        def evalEdges(self):
            self.tap = list(self.EDGES)
            return super().evalEdges()

    # hand-over values a fine acquisition would deliver (nearest 50-Hz bin, integer code phase)
    chans, ours_dec, obs = {}, {}, {}
    for s in sats:
        f = 50.0 * round(con.doppler_at_start(s) / 50.0)
        d = (int(con.code_delay_at_start(s, TOW0, BIAS)) + 1) % 2048
        chans[s.prn] = Tap(s.prn, f, delay=d, itSweep=40, corrMin=8, corrAvg=8, sweepCorrAvg=4)
        ours_dec[s.prn] = navbits.FrameDecoder()
        obs[s.prn] = pos.ChannelObservables(s.prn, N_CYC)

    coPhLst, ref_fixes, frames_equal = {}, [], True
    for e in range(n_ep):
        im, re = np.divmod(raw[2 * e * ngps:2 * (e + 1) * ngps].view(np.uint16), 256)      # gpsrecv.py:168-173
        data = np.asarray(re + 1j * im, dtype=np.complex64) / 127.5 - (1 + 1j)
        smp = np.int64((e + 1) * ngps)
        frameLst = []
        for prn, ch in chans.items():
            ch.tap = None
            swFq, fLst, coPh, cpQ = ch.process(data, smp)
            frameLst += fLst
            if coPh >= 0:
                coPhLst.setdefault(prn, []).append((int(smp // ngps), coPh))
            obs[prn].add_epoch(int(smp), coPh, float(ch.FREQ))
            if ch.tap is not None and len(ch.tap) > 2:        # our decoder on the EDGES list the reference decoder just saw
                mine = ours_dec[prn](None, ch.tap)
                theirs = [{k: (v.item() if hasattr(v, "item") else v) for k, v in f.items()
                           if k not in ("SAT", "AMP", "CRM", "FRQ", "SWP", "EPH")} for f in fLst if "ID" in f]
                mine = [{k: (v.item() if hasattr(v, "item") else v) for k, v in f.items()} for f in mine]
                frames_equal &= (mine == theirs)
            obs[prn].add_frames([f for f in fLst if "ID" in f])
        if len(frameLst) > 0:                                                    # gpsrecv.py:507-519 -> gpseval.py:529-540
            cpLst, _ = gpseval.prepCodePhase(coPhLst, 0)
            satResLst, _, _, actSats, gpsTime = gpseval.evalData(frameLst, cpLst, {}, {})
            satPosLst, recPosLst, failLst = gpseval.ecefPositions(satResLst, None)
            for satNo in coPhLst:
                gpseval.COPH_LIST[satNo] = gpseval.COPH_LIST.get(satNo, []) + coPhLst[satNo]
            ref_fixes += [[float(v) for v in p] for p in recPosLst]
            coPhLst = {}

    ref_xyz = np.array([p[1:4] for p in ref_fixes]) if ref_fixes else np.zeros((0, 3))
    ready = [o for o in obs.values() if o.ready]
    s_rx = float((n_ep - 2) * ngps)
    ttx = np.array([o.transmit_time(s_rx) for o in ready])
    p_ours, cb, res = pos.solve_fix([o.eph for o in ready], ttx, s_rx / pos.FS)
    out = {"truth_ecef": rx.tolist(), "n_reference_fixes": int(len(ref_xyz)),
           "reference_fix_mean": ref_xyz.mean(axis=0).tolist() if len(ref_xyz) else None,
           "reference_fix_err_m": [float(np.linalg.norm(p - rx)) for p in ref_xyz],
           "our_fix": p_ours.tolist(), "our_fix_err_m": float(np.linalg.norm(p_ours - rx)),
           "ours_vs_reference_mean_m": float(np.linalg.norm(p_ours - ref_xyz.mean(axis=0))) if len(ref_xyz) else None,
           "our_frames_equal_reference_frames": bool(frames_equal), "channels_ready": len(ready)}
    print(json.dumps(out, indent=1))
    with open(os.path.join(ROOT, "tests", "golden", "e2e_reference_fix.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
