"""Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE (build container only).

    python oracle/make_golden.py            # runs itself once per N_CYC value

/root/reference/src is put on sys.path (with oracle/_stubs/rtlsdr.py standing in
for the absent RTL-SDR driver) and the unmodified reference
functions are executed on seeded synthetic uint8 I/Q.  While doing so the
numpy restatement in oracle/gps_oracle.py is run on the same inputs and every
output is asserted to be BIT-IDENTICAL to the reference's; only then are the
reference's outputs written as fixtures.  The fixtures (plus this script) are
what travels to the GPU box; /root/reference does not.

N_CYC is an import-time constant of the reference (gpsglob.py:122-125), so each
value needs a fresh interpreter: the parent re-invokes this file with --ncyc.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF_SRC = "/root/reference/src"


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _eq(a, b, what):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype or not np.array_equal(a, b, equal_nan=True):
        raise AssertionError(f"oracle != reference for {what}: {a!r} vs {b!r}")


def _scalar_eq(a, b, what):
    if type(a) is not type(b) and not (isinstance(a, (int, float)) and isinstance(b, (int, float))):
        raise AssertionError(f"type mismatch for {what}: {type(a)} vs {type(b)}")
    if not (a == b or (a != a and b != b)):
        raise AssertionError(f"oracle != reference for {what}: {a!r} vs {b!r}")


def scenario(n_cyc: int):
    """Synthetic recording shared by the generator and the tests."""
    sys.path.insert(0, ROOT)
    from gps_sdr_receiver_b200 import synth
    sats = synth.default_constellation(6, seed=5)
    n_epochs = 72 if n_cyc == 32 else 264
    return synth, sats, n_epochs


def run(n_cyc: int):
    sys.path.insert(0, REF_SRC)
    sys.path.insert(0, os.path.join(HERE, "_stubs"))
    sys.path.insert(0, ROOT)
    import gpsglob
    gpsglob.N_CYC = n_cyc
    gpsglob.NGPS = n_cyc * gpsglob.CODE_SAMPLES
    import gpslib
    import gpsrecv
    from cacodes import cacodes
    from scipy.fft import fft, ifft
    from oracle import gps_oracle as orc

    gpsrecv.FFT_CACODE = [0, 0] + [fft(gpslib.GPSCacode(p)) for p in gpsrecv.SAT_ALL]
    ngps = gpsglob.NGPS
    synth, sats, n_epochs = scenario(n_cyc)
    raw = synth.make_iq(sats, n_cyc * n_epochs, noise_sigma=0.25, seed=11)
    out = {"n_cyc": n_cyc, "raw_sha": _sha(raw), "n_epochs": n_epochs,
           "sat_prn": [s.prn for s in sats], "sat_doppler": [s.doppler for s in sats],
           "sat_delay": [s.delay for s in sats]}

    def block(e):
        return raw[e * 2 * ngps:(e + 1) * 2 * ngps]

    def ref_conv(rb):          # gpsrecv.py:168-173, verbatim semantics of the reader
        im, re = np.divmod(rb.view(np.uint16), 256)
        return np.asarray(re + 1j * im, dtype=np.complex64) / 127.5 - (1 + 1j)

    # ---- (1) tables -------------------------------------------------------
    if n_cyc == 32:
        chips = np.stack([np.asarray(cacodes[p], dtype=np.int8) for p in range(1, 38)])
        for p in range(1, 38):
            _eq(orc.ca_chips(p), chips[p - 1], f"chips prn {p}")
        codes = np.stack([gpslib.GPSCacode(p) for p in range(1, 38)])
        for p in range(1, 38):
            _eq(orc.ca_code_2048(p), codes[p - 1], f"GPSCacode prn {p}")
            _eq(orc.code_spectrum(p), fft(gpslib.GPSCacode(p)), f"spectrum prn {p}")
        assert np.array_equal(codes.astype(np.float32).astype(np.float64), codes)
        spec = np.stack([fft(codes[p - 1]) for p in range(1, 38)])
        np.savez_compressed(os.path.join(GOLD, "tables.npz"),
                            chips_packed=np.packbits(chips > 0, axis=1),
                            code_f32=codes.astype(np.float32),       # exact (asserted above)
                            code_sha_1_32=_sha(codes[:32]),
                            spectrum_prn1_7_19=spec[[0, 6, 18]],
                            spectrum_sha=_sha(spec))
        _eq(orc.raw_to_complex(block(0)), ref_conv(block(0)), "raw->complex")
        _eq(orc.sec_time(ngps), gpsrecv.SEC_TIME, "SEC_TIME")
        print("tables ok; code sha", _sha(codes[:32])[:16])

    # ---- (2) cold-start sweep over 5 streams --------------------------------
    freq_r = freq_o = gpsglob.MIN_FREQ
    lst_r, lst_o = gpsrecv.SAT_ALL.copy(), gpsrecv.SAT_ALL.copy()
    found_r, found_o = [], []
    spectra = {p: orc.code_spectrum(p) for p in range(1, 33)}
    sweep_log = []
    e = 0
    ready_r = False
    while not ready_r:
        d = ref_conv(block(e))
        ready_r, freq_r, found_r = gpsrecv.sweepAllSats(d, freq_r, lst_r, found_r, itSweep=gpsglob.IT_SWEEP_ALL)
        ready_o, freq_o, found_o = orc.sweep_all_sats(orc.raw_to_complex(block(e)), freq_o, lst_o, found_o,
                                                      spectra, it_sweep=orc.IT_SWEEP_ALL, n_cyc=n_cyc)
        assert ready_r == ready_o and freq_r == freq_o and lst_r == lst_o
        assert len(found_r) == len(found_o)
        for a, b in zip(found_r, found_o):
            assert a == b, (a, b)
        sweep_log.append((e, ready_r, freq_r, len(found_r)))
        e += 1
    out["sweep_streams"] = e
    out["sweep_found"] = np.array([(z, p, f, d) for z, p, f, d in found_r], dtype=np.float64)
    out["sweep_log"] = np.array(sweep_log, dtype=np.float64)
    print("sweep ok:", [(int(p), f, int(d), round(float(z), 2)) for z, p, f, d in found_r])

    # full z / argmax grid of the first stream (all 31 PRNs x first 10 bins) for magnitude parity
    d0 = ref_conv(block(0))
    zgrid = np.zeros((31, 10))
    mxgrid = np.zeros((31, 10), dtype=np.int64)
    for b in range(10):
        f = gpsglob.MIN_FREQ + b * gpsglob.STEP_FREQ
        nd, _ = gpsrecv.demodDoppler(d0, f, 0, 4 * 2048)
        sp = sum(fft(nd[i * 2048:(i + 1) * 2048]) for i in range(4)) / 4
        for i, p in enumerate(gpsrecv.SAT_ALL):
            c = np.abs(ifft(sp * np.conjugate(gpsrecv.FFT_CACODE[p])))
            dl, z = gpsrecv.findCodePhase(c)
            zgrid[i, b] = z
            mxgrid[i, b] = np.argmax(c)
    out["sweep_zgrid"] = zgrid
    out["sweep_mxgrid"] = mxgrid

    # ---- (3)-(5) tracking trajectories ------------------------------------------
    start_e = e                                  # tracking starts on the stream after the sweep
    chans = [(int(p), float(f), int(dl)) for _, p, f, dl in found_r]
    out["chan_init"] = np.array(chans, dtype=np.float64)
    force_sweep_at = {2: start_e + 40} if n_cyc == 32 else {2: start_e + 150}   # channel idx -> epoch
    gap_at = start_e + 30 if n_cyc == 32 else start_e + 100                       # one stream dropped
    traj = {}
    for ci, (prn, f0, dl0) in enumerate(chans):
        ref = gpslib.SatStream(prn, f0, delay=dl0, itSweep=gpsglob.IT_SWEEP, corrMin=gpsglob.CORR_MIN,
                               corrAvg=gpsglob.CORR_AVG, sweepCorrAvg=gpsglob.SWEEP_CORR_AVG)
        ref.CALC_PLOT = True
        och = orc.Channel(prn, f0, delay=dl0, n_cyc=n_cyc)
        rows, prompts, edges, corr_keep = [], [], [], {}
        smp = np.int64(start_e) * ngps
        for ep in range(start_e, n_epochs):
            smp = smp + ngps
            if ep == gap_at:                 # this stream is dropped (ring-buffer skip, gpsrecv.py:469-471)
                continue
            force = force_sweep_at.get(ci) == ep
            d = ref_conv(block(ep))
            took_track = not ref.SWEEP and not force
            sw_r, frames, cp_r, (q_r, l_r) = ref.process(d, smp, sweep=force)
            sw_o, rep_o, cp_o, (q_o, l_o) = och.process(orc.raw_to_complex(block(ep)), smp, sweep=force)
            # -- bit-exact oracle == reference ---------------------------------
            assert sw_r == sw_o and (len(frames) > 0) == rep_o, (ep, sw_r, sw_o, frames, rep_o)
            for a, b, w in ((cp_r, cp_o, "codePhase"), (q_r, q_o, "corrQ"), (l_r, l_o, "corrL"),
                            (ref.FREQ, och.freq, "FREQ"), (ref.PHASE, och.phase, "PHASE"),
                            (ref.DELAY, och.delay, "DELAY"), (ref.MAX_CORR, och.max_corr, "MAX_CORR"),
                            (ref.AMPLITUDE, och.amplitude, "AMP"), (ref.STD_DEV, och.std_dev, "STD"),
                            (ref.MS_TIME, och.ms_time, "MS_TIME"), (ref.PHASE_LOCKED, och.locked, "LOCKED")):
                _scalar_eq(a, b, f"{w} prn {prn} epoch {ep}")
            assert ref.EDGES == och.edges, (ep, ref.EDGES, och.edges)
            assert [float(v) for v in ref.DF] == [float(v) for v in och.df]
            assert len(ref.PREV_SAMPLES) == len(och.prev_samples)
            if took_track:                    # prompts only exist on the tracking branch
                _eq(np.asarray(ref.GPSDATA), och.prompt, f"gpsData prn {prn} epoch {ep}")
            rows.append([float(sw_r), float(cp_r), float(q_r), float(l_r), float(ref.FREQ), float(ref.PHASE),
                         float(ref.DELAY), float(ref.MAX_CORR), float(ref.AMPLITUDE), float(ref.STD_DEV),
                         float(ref.MS_TIME), float(ref.PHASE_LOCKED), float(len(frames) > 0),
                         float(len(ref.PREV_SAMPLES)), float(ep), float(smp), float(force),
                         float(frames[0].get("SWP", False)) if frames else 0.0, float(took_track)])
            prompts.append(np.asarray(och.prompt) if took_track else np.zeros(0, dtype=np.complex64))
            edges.append(list(och.new_edges))
            if took_track and ep in (start_e, start_e + 5):
                corr_keep[ep] = och.last_corr.copy()
        pr_len = np.array([len(p) for p in prompts])
        pr = np.zeros((len(prompts), n_cyc + 2), dtype=np.complex64)
        for i, p in enumerate(prompts):
            pr[i, :len(p)] = p
        ed = [(i, ms, st) for i, lst in enumerate(edges) for ms, st in lst]
        traj[f"ch{ci}_rows"] = np.array(rows)
        traj[f"ch{ci}_prompt"] = pr
        traj[f"ch{ci}_prompt_len"] = pr_len
        traj[f"ch{ci}_edges"] = np.array(ed, dtype=np.int64).reshape(-1, 3)
        for ep, c in corr_keep.items():
            traj[f"ch{ci}_corr_e{ep - start_e}"] = c
        locked_at = next((i for i, r in enumerate(rows) if r[11] > 0), -1)
        print(f"  prn {prn}: rows {len(rows)} locked at row {locked_at} edges {len(ed)} "
              f"final f {rows[-1][4]:.2f} cp {rows[-1][1]:.3f} swept {sum(r[0] for r in rows):.0f}")
    out.update(traj)
    out["start_epoch"] = start_e
    out["gap_at"] = gap_at
    out["row_cols"] = np.array(["sweep", "codePhase", "corrQ", "corrL", "FREQ", "PHASE", "DELAY", "MAX_CORR",
                                "AMPLITUDE", "STD_DEV", "MS_TIME", "LOCKED", "report", "n_prev", "epoch",
                                "smpTime", "forced", "SWP", "tracked"])

    # ---- decodeData at fixed delays on a fresh channel ---------------------------
    if n_cyc == 32:
        prn, f0, dl0 = chans[0]
        for dl in (0, 1, 417, 2047):
            ref = gpslib.SatStream(prn, f0, delay=dl)
            ref.SMP_TIME = np.int64(ngps)
            och = orc.Channel(prn, f0, delay=dl, n_cyc=n_cyc)
            och.smp_time = np.int64(ngps)
            d = ref_conv(block(start_e))
            g1 = ref.decodeData(d, dl)
            g2 = och._decode(orc.raw_to_complex(block(start_e)), dl)
            _eq(g1, g2, f"decodeData delay {dl}")
            d = ref_conv(block(start_e + 1))
            dl2 = (dl + 2047) % 2048 if dl else 3     # move the delay backwards / across the wrap
            g1b = ref.decodeData(d, dl2)
            g2b = och._decode(orc.raw_to_complex(block(start_e + 1)), dl2)
            _eq(g1b, g2b, f"decodeData delay {dl}->{dl2}")
            out[f"decode_d{dl}_a"] = g1
            out[f"decode_d{dl}_b"] = g1b
            out[f"decode_d{dl}_next"] = dl2

    # ---- generalised grid (reference-derived oracle; small case) -------------------
    if n_cyc == 32:
        data = orc.raw_to_complex(raw[:2 * 10 * 2048])
        prns = [sats[0].prn, sats[2].prn, 1, 32]
        g = orc.acq_grid(data, prns, -1000.0, 500.0, 5, 1, 10, orc.ACQ_MODE_POW)
        # cross-check the ABS mode of the grid against the reference's own sweep primitives
        g_abs = orc.acq_grid(ref_conv(block(0))[:4 * 2048], gpsrecv.SAT_ALL, -5000.0, 200.0, 10, 4, 1,
                             orc.ACQ_MODE_ABS)
        assert np.array_equal(g_abs["mx"], mxgrid) and np.array_equal(g_abs["z"], zgrid)
        for k, v in g.items():
            out[f"grid_{k}"] = v
        out["grid_prns"] = np.array(prns)

    np.savez_compressed(os.path.join(GOLD, f"traj_ncyc{n_cyc}.npz"), **out)
    print(f"wrote traj_ncyc{n_cyc}.npz")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncyc", type=int, default=0)
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    if a.ncyc:
        run(a.ncyc)
    else:
        for n in (32, 8):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--ncyc", str(n)])
