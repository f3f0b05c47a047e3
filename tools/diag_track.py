"""GPU diagnostic: per-epoch differences between the tracker kernel and the golden trajectories."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest
from gps_sdr_receiver_b200.tracking import TrackBank, new_edges

n_cyc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
chs = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2]
scen = conftest.Scenario(n_cyc)
g = scen.gold
start_e, gap_at = int(g["start_epoch"]), int(g["gap_at"])
ngps = scen.ngps
bank = TrackBank(n_cyc, 8)
slots = [bank.add(int(p), float(f), int(d)) for p, f, d in g["chan_init"]]
FORCED = 16
forced = {ci: int(g[f"ch{ci}_rows"][g[f"ch{ci}_rows"][:, FORCED] > 0][0, 14]) for ci in range(6) if (g[f"ch{ci}_rows"][:, FORCED] > 0).any()}
cuts = sorted(set([start_e, gap_at, gap_at + 1, scen.n_epochs] + list(forced.values())))
out = []
for a, b in zip(cuts[:-1], cuts[1:]):
    if a == gap_at:
        continue
    for ci, ep in forced.items():
        if ep == a:
            bank.request_sweep(slots[ci])
    out.append(bank.process(scen.raw[a * 2 * ngps:b * 2 * ngps], (a + 1) * ngps, b - a))
recs = np.concatenate(out, axis=0)
n_ep = len(recs)
for ci in chs:
    rows = g[f"ch{ci}_rows"]
    print("channel", ci, g["chan_init"][ci])
    eg = g[f"ch{ci}_edges"]
    got = [(r, ms, st) for r in range(n_ep) for ms, st in new_edges(recs[r, ci])]
    ref = [tuple(x) for x in eg if x[0] < n_ep]
    print(" edges got", len(got), "ref", len(ref), "only-got", sorted(set(got) - set(ref))[:6], "only-ref", sorted(set(ref) - set(got))[:6])
    for r in range(n_ep):
        rec = recs[r, ci]
        n = int(g[f"ch{ci}_prompt_len"][r])
        pg = np.ascontiguousarray(rec["prompt"][:2 * n]).view(np.complex64)
        pr = g[f"ch{ci}_prompt"][r][:n]
        dd = np.abs(pg - pr) if n else np.zeros(1)
        pr = pr if n else np.ones(1)
        if True:
            print(f" r{r:3d} dfreq {rec['freq'] - rows[r][4]:+.3e} dphase {rec['phase'] - rows[r][5]:+.3e} dcp {rec['code_phase'] - rows[r][1]:+.2e} "
                  f"dz {rec['max_corr'] / rows[r][7] - 1:+.1e} lock {rec['locked']}/{int(rows[r][11])} prompt d0 {dd[0]:.1e} dmax {dd.max():.1e} rel {dd.max() / np.abs(pr).max():.1e} "
                  f"std {rec['std_dev'] / rows[r][9] - 1:+.1e}")
