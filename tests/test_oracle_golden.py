"""The oracle against the golden vectors recorded from the REAL reference
(oracle/make_golden.py ran /root/reference/src in the build container and stored its
outputs).  CPU only.  If these pass, the oracle restates the reference bit for bit on
this machine's numpy/scipy as well."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import pytest

from oracle import gps_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_tables_match_reference():
    g = np.load(os.path.join(GOLD, "tables.npz"))
    chips = np.unpackbits(g["chips_packed"], axis=1)[:, :1023].astype(np.int8) * 2 - 1
    codes = np.stack([orc.ca_code_2048(p) for p in range(1, 38)])
    for p in range(1, 38):
        assert np.array_equal(orc.ca_chips(p), chips[p - 1])
    assert np.array_equal(codes, g["code_f32"].astype(np.float64))
    assert _sha(codes[:32]) == str(g["code_sha_1_32"])
    assert _sha(codes[:32]).startswith("3c7b151519bfac8b")
    # SURVEY.md 8c known answers for PRN 1
    c1 = codes[0]
    assert c1.sum() == 2.001169204711914
    assert int((np.abs(c1) == 1).sum()) == 1537
    spec = np.stack([orc.code_spectrum(p) for p in (1, 7, 19)])
    np.testing.assert_allclose(spec, g["spectrum_prn1_7_19"], rtol=0, atol=1e-9)
    assert abs(spec[0][1] - (-31.06851935586514 + 31.59978701067999j)) < 1e-9
    assert np.array_equal(orc.ca_chips(34), orc.ca_chips(37))        # PRN34 == PRN37 in the reference table


def _run_sweep(scen):
    freq = orc.MIN_FREQ
    lst = list(range(2, 33))
    found = []
    spectra = {p: orc.code_spectrum(p) for p in range(1, 33)}
    log = []
    e, ready = 0, False
    while not ready:
        ready, freq, found = orc.sweep_all_sats(orc.raw_to_complex(scen.block(e)), freq, lst, found, spectra,
                                                it_sweep=orc.IT_SWEEP_ALL, n_cyc=scen.n_cyc)
        log.append((e, ready, freq, len(found)))
        e += 1
    return e, found, log


@pytest.mark.parametrize("which", ["scen32", "scen8"])
def test_cold_start_sweep_matches_reference(which, request):
    scen = request.getfixturevalue(which)
    e, found, log = _run_sweep(scen)
    g = scen.gold
    assert e == int(g["sweep_streams"])
    got = np.array([(z, p, f, d) for z, p, f, d in found], dtype=np.float64)
    assert np.array_equal(got, g["sweep_found"])
    assert np.array_equal(np.array(log, dtype=np.float64), g["sweep_log"])


@pytest.mark.parametrize("which", ["scen32", "scen8"])
def test_tracking_trajectories_match_reference(which, request):
    scen = request.getfixturevalue(which)
    g = scen.gold
    start_e, gap_at = int(g["start_epoch"]), int(g["gap_at"])
    for ci, (prn, f0, dl0) in enumerate(g["chan_init"]):
        rows_g = g[f"ch{ci}_rows"]
        ch = orc.Channel(int(prn), float(f0), delay=int(dl0), n_cyc=scen.n_cyc)
        smp = np.int64(start_e) * scen.ngps
        r = 0
        edges = []
        for ep in range(start_e, scen.n_epochs):
            smp = smp + scen.ngps
            if ep == gap_at:
                continue
            force = bool(rows_g[r][16])
            sw, rep, cp, (q, l) = ch.process(orc.raw_to_complex(scen.block(ep)), smp, sweep=force)
            row = [float(sw), float(cp), float(q), float(l), float(ch.freq), float(ch.phase), float(ch.delay),
                   float(ch.max_corr), float(ch.amplitude), float(ch.std_dev), float(ch.ms_time), float(ch.locked),
                   float(rep), float(len(ch.prev_samples)), float(ep), float(smp), float(force)]
            assert row == list(rows_g[r][:17]), (ci, ep, row, rows_g[r])
            if rows_g[r][18]:
                n = int(g[f"ch{ci}_prompt_len"][r])
                assert np.array_equal(ch.prompt, g[f"ch{ci}_prompt"][r][:n])
            edges += [(r, ms, st) for ms, st in ch.new_edges]
            r += 1
        assert r == len(rows_g)
        assert np.array_equal(np.array(edges, dtype=np.int64).reshape(-1, 3), g[f"ch{ci}_edges"])


def test_decode_fixed_delays_match_reference(scen32):
    g = scen32.gold
    start_e = int(g["start_epoch"])
    prn, f0, _ = g["chan_init"][0]
    for dl in (0, 1, 417, 2047):
        ch = orc.Channel(int(prn), float(f0), delay=dl, n_cyc=32)
        ch.smp_time = np.int64(scen32.ngps)
        a = ch._decode(orc.raw_to_complex(scen32.block(start_e)), dl)
        assert np.array_equal(a, g[f"decode_d{dl}_a"])
        b = ch._decode(orc.raw_to_complex(scen32.block(start_e + 1)), int(g[f"decode_d{dl}_next"]))
        assert np.array_equal(b, g[f"decode_d{dl}_b"])


def test_generalised_grid_fixture(scen32):
    """acq_grid (reference-derived) against its committed fixture, and its ABS mode
    against the reference's own sweep primitives (sweep_zgrid / sweep_mxgrid)."""
    g = scen32.gold
    data = orc.raw_to_complex(scen32.raw[:2 * 10 * 2048])
    out = orc.acq_grid(data, [int(p) for p in g["grid_prns"]], -1000.0, 500.0, 5, 1, 10, orc.ACQ_MODE_POW)
    for k, v in out.items():
        assert np.array_equal(v, g[f"grid_{k}"]), k
    g_abs = orc.acq_grid(orc.raw_to_complex(scen32.block(0))[:4 * 2048], list(range(2, 33)), -5000.0, 200.0, 10, 4, 1,
                         orc.ACQ_MODE_ABS)
    assert np.array_equal(g_abs["mx"], g["sweep_mxgrid"])
    assert np.array_equal(g_abs["z"], g["sweep_zgrid"])
