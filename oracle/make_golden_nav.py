"""Golden vectors for the nav-bit decoder: RUNS THE REAL REFERENCE (build container only).

    python oracle/make_golden_nav.py

EDGES lists (what the tracker hands to evalEdges once per second) are synthesised from LNAV bit
streams made by gps_sdr_receiver_b200.navbits.encode_frames -- polarity inverted, with a bit error,
with a preamble look-alike in the data, cut into once-per-second calls -- and fed to the unmodified
`gpslib.SatStream.evalEdges` (src/gpslib.py:1451-1580).  The frame dicts it returns and the bits it
keeps are stored in tests/golden/navbits.json next to the inputs.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference/src")

EPH = dict(weekNum=345, satAcc=2, satHealth=0, Tgd=-1.1175870895385742e-08, IODC=0x2A5, Toc=230400, af2=0.0,
           af1=-3.637978807091713e-12, af0=-0.00021193176507949829,
           IODE2=0xA5, Crs=-88.65625, deltaN=4.3e-09, M0=-1.2345678, Cuc=-4.5e-06, e=0.0123456789, Cus=8.1e-06,
           sqrtA=5153.6789, Toe=230400,
           Cic=-1.1e-07, omegaBig=2.3456789, Cis=9.5e-08, i0=0.9765432, IODE3=0xA5, Crc=250.25, omegaSmall=-0.87654321,
           omegaDot=-8.1e-09, IDOT=2.5e-10)


def bits_to_edges(bits01: np.ndarray, ms0: int, st0: int, first_sign: int):
    """+-1 bit stream -> EDGES list the tracker would have built: [sign before the first edge,
    (ms, sample time) of every sign change], 20 ms per bit, 2048 samples per ms."""
    pm = np.where(bits01 > 0, 1, -1) * first_sign
    edges = [int(-pm[0])]                      # the signal before the first recorded edge had the opposite sign
    prev = -pm[0]
    for k, b in enumerate(pm):
        if b != prev:
            ms = ms0 + 20 * k
            edges.append((int(ms), int(st0 + 2048 * (ms - ms0))))
            prev = b
    return edges


def cases():
    from gps_sdr_receiver_b200 import navbits
    rng = np.random.default_rng(7)
    out = []
    base = navbits.encode_frames(100, 7, EPH, first_id=4)          # ids 4 5 1 2 3 4 5
    lead = rng.integers(0, 2, 37).astype(np.int8)
    tail = rng.integers(0, 2, 23).astype(np.int8)
    stream = np.concatenate([lead, base, tail])
    out.append(("plain", stream, 1, 1))
    out.append(("inverted", stream, -1, 1))
    bad = stream.copy()
    bad[37 + 300 * 2 + 95] ^= 1                                   # one bit error in subframe id 1
    out.append(("bit_error", bad, 1, 1))
    fake = stream.copy()
    fake[37 + 300 * 3 + 100:37 + 300 * 3 + 108] = [1, 0, 0, 0, 1, 0, 1, 1]   # preamble look-alike inside a data word
    out.append(("lookalike", fake, 1, 1))
    out.append(("chunked", stream, 1, 9))                         # delivered in 9 calls, like once per second
    short = np.concatenate([lead, base[:280]])
    out.append(("too_short", short, 1, 1))
    return out


def main():
    import gpslib                                                 # the reference
    res = []
    for name, bits, sign, n_calls in cases():
        edges = bits_to_edges(bits, 1000, 5_000_000, sign)
        ch = gpslib.SatStream(7, 0.0)
        body = edges[1:]
        cuts = np.linspace(0, len(body), n_calls + 1).astype(int)
        calls, frames_per_call = [], []
        ch.EDGES = [edges[0]]
        for c in range(n_calls):
            ch.EDGES = ch.EDGES + body[cuts[c]:cuts[c + 1]]
            calls.append([edges[0] if c == 0 else None, body[cuts[c]:cuts[c + 1]]])
            fr = ch.evalEdges()
            frames_per_call.append([{k: (v.item() if hasattr(v, "item") else v) for k, v in f.items()} for f in fr])
        res.append({"name": name, "first_sign": edges[0], "edge_chunks": [c[1] for c in calls], "frames": frames_per_call,
                    "kept_bits": [int(b) for b in ch.GPSBITS], "kept_edges": [ch.EDGES[0]] + [list(e) for e in ch.EDGES[1:]]})
        print(name, [len(f) for f in frames_per_call], "kept", len(ch.GPSBITS))
    with open(os.path.join(ROOT, "tests", "golden", "navbits.json"), "w") as f:
        json.dump({"eph": EPH, "cases": res}, f)


if __name__ == "__main__":
    main()
