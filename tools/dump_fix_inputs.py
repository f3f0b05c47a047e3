#!/usr/bin/env python
"""GPU box: the once-per-second messages (skippedData, frameLst, coPhLst) that gpsrecv would send to gpseval
(src/gpsrecv.py:496-519) for a geometry-consistent synthetic recording, produced (a) by the GPU hot path and (b) by the
CPU oracle (bit-exact restatement of the reference's SatStream) on the SAME bytes.  Written to
gpurun_out/fix_inputs.json for oracle/e2e_consumer_on_gpu_outputs.py, which feeds both to the unmodified gpseval."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

N_CYC, SECONDS, TOW0, BIAS = 32, 30, 345597, 1.2345e-4


def clean(o):
    if isinstance(o, dict):
        return {str(k): clean(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [clean(v) for v in o]
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.bool_,)):
        return bool(o)
    return o


def main():
    import torch
    from gps_sdr_receiver_b200 import constellation as con, glob, navbits
    from gps_sdr_receiver_b200.acquisition import AcqPlan, GR_ACQ_POW
    from gps_sdr_receiver_b200.tracking import SatStream, TrackBank
    from oracle import gps_oracle as orc
    glob.set_n_cyc(N_CYC)
    rx, sats = con.build(seconds=SECONDS, n_sat=7, tow0=TOW0, rx_clock_bias=BIAS, seed=1)
    n_ms = SECONDS * 1000 // N_CYC * N_CYC
    d_raw = con.make_iq_dev(sats, n_ms, TOW0, BIAS, noise_sigma=0.25, seed=11)
    ngps, n_ep = N_CYC * 2048, n_ms // N_CYC
    no_sec = 1024 // N_CYC
    prns = [s.prn for s in sats]
    bins = [-5000.0 + 50.0 * b for b in range(201)]
    best = AcqPlan.best_from_tensor(AcqPlan(prns, bins, 10, 2, GR_ACQ_POW).search_dev(d_raw))[0]
    init = [(s.prn, bins[int(b["bin"])], int(b["cell"]["mx"])) for s, b in zip(sats, best)]

    # (a) GPU
    bank = TrackBank(N_CYC, 8)
    streams = [SatStream(p, f, delay=d, bank=bank, frame_decoder=navbits.FrameDecoder()) for p, f, d in init]
    recs = TrackBank.records_from_tensor(bank.process_dev(d_raw, ngps, n_ep))
    msgs_gpu, coph = [], {}
    for e in range(n_ep):
        smp, frame_lst = (e + 1) * ngps, []
        for c, st in enumerate(streams):
            _, f_lst, cp, _ = st.absorb(recs[e, c], smp)
            frame_lst += f_lst
            if cp >= 0:
                coph.setdefault(st.SAT_NO, []).append((smp // ngps, float(cp)))
        if frame_lst:
            msgs_gpu.append((0, frame_lst, coph))
            coph = {}
    for st in streams:
        st.close()
    bank.close()

    # (b) oracle on the same bytes
    raw = d_raw.cpu().numpy()
    class DecodingChannel(orc.Channel):               # evalEdges = decoder on the full EDGES list, then logicalBits' trim
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.decoder, self.last_frames = navbits.FrameDecoder(), []

        def consume_edges(self):
            if len(self.edges) > 2:
                self.last_frames = self.decoder(None, list(self.edges))
            super().consume_edges()

    chans = [DecodingChannel(p, f, delay=d, n_cyc=N_CYC) for p, f, d in init]
    msgs_cpu, coph = [], {}
    for e in range(n_ep):
        smp, frame_lst = (e + 1) * ngps, []
        data = orc.raw_to_complex(raw[2 * e * ngps:2 * (e + 1) * ngps])
        for ch, (p, _, _) in zip(chans, init):
            ch.last_frames = []
            _, report_due, cp, _ = ch.process(data, np.int64(smp))
            if cp >= 0:
                coph.setdefault(p, []).append((smp // ngps, float(cp)))
            if report_due:                                  # frameLst or [{}], then reportValues (gpslib.py:1191-1197, 1124-1131)
                frames = [dict(f) for f in ch.last_frames] or [{}]
                for f in frames:
                    f.update(SAT=p, AMP=float(ch.amplitude), CRM=float(ch.max_corr), FRQ=float(ch.freq), SWP=bool(ch.rep_sweep_reported))
                frame_lst += frames
        if frame_lst:
            msgs_cpu.append((0, frame_lst, coph))
            coph = {}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fix_inputs.json"), "w") as f:
        json.dump(clean({"truth_ecef": rx.tolist(), "n_cyc": N_CYC, "init": init, "gpu": msgs_gpu, "cpu": msgs_cpu}), f)
    same_frames = sum(1 for a, b in zip(msgs_gpu, msgs_cpu) for fa, fb in zip(a[1], b[1])
                      if {k: v for k, v in fa.items() if k in ("ID", "tow", "ST", "SAT")} == {k: v for k, v in fb.items() if k in ("ID", "tow", "ST", "SAT")})
    print("messages", len(msgs_gpu), len(msgs_cpu), "frame dicts with equal (SAT, ID, tow, ST):", same_frames, "of", sum(len(m[1]) for m in msgs_gpu))


if __name__ == "__main__":
    main()
