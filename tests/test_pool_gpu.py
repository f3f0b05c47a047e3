"""The receiver loop of gpsrecv.processData (src/gpsrecv.py:445-548) driven headless on the
B200 path through the drop-in functions (sweepAllSats, initMultiProcPool, initPoolStreams,
satCalc, delPoolStreams), against the oracle run through the same state machine."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import gps_oracle as orc

pytestmark = pytest.mark.gpu


def test_cold_start_then_tracking_like_processData(gpu, scen32):
    from gps_sdr_receiver_b200 import glob, pool as gp
    from gps_sdr_receiver_b200.acquisition import sweepAllSats
    glob.set_n_cyc(32)
    ngps = scen32.ngps
    # ---- sweep-all state (gpsrecv.py:468-490) ----
    freq, satLst, foundSats = glob.MIN_FREQ, list(range(2, 33)), []
    e, ready = 0, False
    while not ready:
        ready, freq, foundSats = sweepAllSats(orc.raw_to_complex(scen32.block(e)), freq, satLst, foundSats, itSweep=glob.IT_SWEEP_ALL)
        e += 1
    assert [int(p) for _, p, _, _ in foundSats] == [int(p) for p in scen32.gold["sweep_found"][:, 1]]
    # ---- pool (gpsrecv.py:453-457, 385-401) ----
    pool, poolNo, poolWorker = gp.initMultiProcPool(glob.MAX_SAT)
    newSatSet = {int(p) for _, p, _, _ in foundSats}
    poolWorker, actSatSet = gp.initPoolStreams(pool, poolNo, poolWorker, set(), set(newSatSet), foundSats)
    assert actSatSet == newSatSet and sorted(w for w in poolWorker if w) == sorted(newSatSet)
    och = {int(p): orc.Channel(int(p), f, delay=int(d), n_cyc=32) for _, p, f, d in foundSats}
    # ---- tracking state (gpsrecv.py:492-519) ----
    coPhLst = {p: [] for p in actSatSet}
    reports = 0
    for ep in range(e, e + 36):
        smp = np.int64((ep + 1) * ngps)
        data = orc.raw_to_complex(scen32.block(ep))
        res = gp.satCalc(actSatSet, pool, poolWorker, data, smp)
        assert len(res) == len(actSatSet)
        for swFq, satNo, frameData, coPh, cpQ in res:
            o = och[satNo]
            sw_o, rep_o, cp_o, (q_o, l_o) = o.process(data, smp)
            assert swFq == sw_o and (len(frameData) > 0) == rep_o and cpQ == (q_o, l_o)
            assert abs(coPh - cp_o) < 2e-4
            if coPh >= 0:
                coPhLst[satNo].append((ep, coPh))
            if frameData:
                reports += 1
                assert frameData[0]["SAT"] == satNo and abs(frameData[0]["FRQ"] - float(o.freq)) < 50.0
    assert reports >= len(actSatSet) and all(len(v) > 25 for v in coPhLst.values())
    # ---- drop two satellites (gpsrecv.py:370-382), the rest keeps running ----
    drop = set(sorted(actSatSet)[:2])
    poolWorker, actSatSet = gp.delPoolStreams(pool, poolNo, poolWorker, actSatSet, drop)
    assert poolWorker.count(0) == poolNo - len(actSatSet) and not (drop & actSatSet)
    ep = e + 36
    smp = np.int64((ep + 1) * ngps)
    data = orc.raw_to_complex(scen32.block(ep))
    res = gp.satCalc(actSatSet, pool, poolWorker, data, smp)
    assert {r[1] for r in res} == actSatSet
    for swFq, satNo, frameData, coPh, cpQ in res:
        _, _, cp_o, _ = och[satNo].process(data, smp)
        assert abs(coPh - cp_o) < 2e-4
    gp.closeMultiProcPool(pool)
