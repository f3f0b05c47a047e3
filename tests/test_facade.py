"""The module INTEGRATION.md tells a maintainer to import instead of the reference's gpslib
(`import gps_sdr_receiver_b200.gpslib as gpslib`, /root/reference/src/gpsrecv.py:4-7, 316-320, 577): hot-path names
come from the CUDA library, everything else falls through to the reference's own module when it is importable."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import pytest

from oracle import gps_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF_SRC = "/root/reference/src"


def _facade():
    return importlib.import_module("gps_sdr_receiver_b200.gpslib")


def test_code_functions_equal_the_reference_tables(built_lib):
    gpslib = _facade()
    g = np.load(os.path.join(GOLD, "tables.npz"))
    for prn in (1, 2, 17, 32, 37):
        code = gpslib.GPSCacode(prn)
        assert code.dtype == np.float64 and code.shape == (2048,)
        assert np.array_equal(code, g["code_f32"][prn - 1].astype(np.float64))
    # GPSCacodeRep(satNo, NCopies, delay) (gpslib.py:81-87): NCopies copies, rolled by delay
    for prn, ncop, delay in ((5, 8, 0), (5, 8, 417), (23, 32, 2047), (9, 1, 3), (9, 16, 2048 + 5)):
        rep = gpslib.GPSCacodeRep(prn, ncop, delay)
        base = g["code_f32"][prn - 1].astype(np.float64)
        y1 = base
        for _ in range(ncop - 1):
            y1 = np.append(y1, base)
        assert rep.dtype == np.float64 and np.array_equal(rep, np.roll(y1, delay))
    # fft(GPSCacode(prn)) is what gpsrecv builds FFT_CACODE from (gpsrecv.py:574-577)
    ref = g["spectrum_prn1_7_19"][1]
    assert np.abs(gpslib.code_spectrum(7) - ref).max() < 1e-9 * np.abs(ref).max()
    assert gpslib.SatStream.__init__.__code__.co_varnames[:8] == (
        "self", "satNo", "freq", "itSweep", "corrMin", "corrAvg", "sweepCorrAvg", "delay")   # gpslib.py:1050-1051


def test_names_off_the_hot_path_fall_through_to_the_reference(built_lib, monkeypatch):
    gpslib = _facade()
    if not os.path.isdir(REF_SRC):
        pytest.skip("reference tree not present (GPU box): covered by the AttributeError test")
    monkeypatch.syspath_prepend(REF_SRC)
    sys.modules.pop("gpslib", None)
    try:
        ref = importlib.import_module("gpslib")
        for name in ("SatOrbit", "leastSquaresPos", "ecefToGeo", "gpsTime", "locDistFromLatLon", "ecefToAzimElev", "Subframe"):
            assert getattr(gpslib, name) is getattr(ref, name), name
        # hot-path names are NOT taken from the reference
        assert gpslib.SatStream is not ref.SatStream and gpslib.GPSCacode is not ref.GPSCacode
        assert np.array_equal(gpslib.GPSCacode(11), ref.GPSCacode(11))
        assert np.array_equal(gpslib.GPSCacodeRep(11, 4, 99), ref.GPSCacodeRep(11, 4, 99))
        with pytest.raises(AttributeError):
            gpslib.noSuchName
    finally:
        for m in ("gpslib", "gpsglob", "cacodes"):
            sys.modules.pop(m, None)


def test_fall_through_raises_attribute_error_without_the_reference(built_lib, monkeypatch):
    gpslib = _facade()
    monkeypatch.setattr(sys, "path", [p for p in sys.path if os.path.abspath(p) != REF_SRC])
    for m in ("gpslib", "gpsglob", "cacodes"):
        monkeypatch.delitem(sys.modules, m, raising=False)
    with pytest.raises(AttributeError, match="outside the B200 hot path"):
        gpslib.SatOrbit
    assert not hasattr(gpslib, "leastSquaresPos")


@pytest.mark.gpu
def test_satstream_through_the_facade_with_the_reference_call(gpu, scen32):
    """gpsrecv.runProc's exact construction (gpsrecv.py:316-320) and calls (gpsrecv.py:331-333), incl. process(data, smpTime,
    sweep=True), reached through the facade; state and return tuples against the oracle channel run side by side."""
    gpslib = _facade()
    from gps_sdr_receiver_b200 import glob
    glob.set_n_cyc(32)
    IT_SWEEP, CORR_MIN, CORR_AVG, SWEEP_CORR_AVG = glob.IT_SWEEP, glob.CORR_MIN, glob.CORR_AVG, glob.SWEEP_CORR_AVG
    g = scen32.gold
    start_e = int(g["start_epoch"])
    satNo, freq, delay = (int(g["chan_init"][2][0]), float(g["chan_init"][2][1]), int(g["chan_init"][2][2]))
    SATPROC = gpslib.SatStream(satNo, freq, delay=delay, itSweep=IT_SWEEP, corrMin=CORR_MIN, corrAvg=CORR_AVG,
                               sweepCorrAvg=SWEEP_CORR_AVG)
    och = orc.Channel(satNo, freq, delay=delay, n_cyc=32)
    assert SATPROC.SAT_NO == satNo
    smp = np.int64(start_e) * scen32.ngps
    for k, ep in enumerate(range(start_e, start_e + 24)):
        smp = smp + scen32.ngps
        data = orc.raw_to_complex(scen32.block(ep))
        force = k == 12
        swFq, frameData, coPh, cpQ = SATPROC.process(data, smp, sweep=True) if force else SATPROC.process(data, smp)
        sw_o, rep_o, cp_o, q_o = och.process(data, smp, sweep=force)
        assert swFq == sw_o and (len(frameData) > 0) == rep_o and tuple(cpQ) == tuple(q_o), (ep, swFq, sw_o)
        assert (coPh >= 0) == (cp_o >= 0) and abs(coPh - cp_o) < 2e-4
        assert SATPROC.DELAY == och.delay and SATPROC.PHASE_LOCKED == och.locked and SATPROC.MS_TIME == och.ms_time
        assert abs(float(SATPROC.FREQ) - float(och.freq)) <= 1e-6 * abs(float(och.freq)) + 1e-3
    del SATPROC            # `del inst` in gpsrecv.runProc (gpsrecv.py:323-328)


@pytest.mark.gpu
def test_pool_functions_through_the_facade_package(gpu, scen32):
    """initMultiProcPool / initPoolStreams / satCalc / delPoolStreams / closeMultiProcPool (gpsrecv.py:340-417) as
    INTEGRATION.md wires them, two satellites, against per-channel oracle runs."""
    from gps_sdr_receiver_b200 import glob, pool as gpool
    glob.set_n_cyc(32)
    g = scen32.gold
    start_e = int(g["start_epoch"])
    found = [(20.0 + i, int(p), float(f), int(d)) for i, (p, f, d) in enumerate(g["chan_init"][:2])]
    pool, poolNo, poolWorker = gpool.initMultiProcPool(4)
    poolWorker, act = gpool.initPoolStreams(pool, poolNo, poolWorker, set(), {e[1] for e in found}, found)
    assert act == {e[1] for e in found} and sorted(w for w in poolWorker if w) == sorted(act)
    ochs = {prn: orc.Channel(prn, f, delay=d, n_cyc=32) for _, prn, f, d in found}
    smp = np.int64(start_e) * scen32.ngps
    for ep in range(start_e, start_e + 10):
        smp = smp + scen32.ngps
        data = orc.raw_to_complex(scen32.block(ep))
        res = gpool.satCalc(act, pool, poolWorker, data, smp)
        assert [r[1] for r in res] == list(act)
        for swFq, satNo, frameData, coPh, cpQ in res:
            sw_o, rep_o, cp_o, q_o = ochs[satNo].process(data, smp)
            assert swFq == sw_o and tuple(cpQ) == tuple(q_o) and abs(coPh - cp_o) < 2e-4 and (len(frameData) > 0) == rep_o
    gone = {found[0][1]}
    poolWorker, act = gpool.delPoolStreams(pool, poolNo, poolWorker, act, gone)
    assert act == {found[1][1]} and poolWorker.count(0) == 3
    res = gpool.satCalc(act, pool, poolWorker, orc.raw_to_complex(scen32.block(start_e + 10)), smp + scen32.ngps)
    assert len(res) == 1 and res[0][1] == found[1][1]
    gpool.closeMultiProcPool(pool)
