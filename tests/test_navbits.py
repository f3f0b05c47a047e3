"""Nav-bit decoder / LNAV encoder against golden vectors recorded from the
reference's own SatStream.evalEdges (oracle/make_golden_nav.py ran the unmodified src/gpslib.py)."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

from gps_sdr_receiver_b200 import navbits

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "navbits.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def test_parity_is_linear_and_tail_solver_zeroes_d29_d30():
    rng = np.random.default_rng(1)
    for _ in range(200):
        a, b = int(rng.integers(0, 1 << 24)), int(rng.integers(0, 1 << 24))
        assert navbits.word_parity(a ^ b, 0, 0) == navbits.word_parity(a, 0, 0) ^ navbits.word_parity(b, 0, 0)
        p29, p30 = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        d24 = navbits._solve_tail(a >> 2, p29, p30)
        assert d24 >> 2 == a >> 2 and navbits.word_parity(d24, p29, p30) & 3 == 0


def test_every_case_decodes_like_the_reference(gold):
    for case in gold["cases"]:
        dec = navbits.FrameDecoder()
        edges = [case["first_sign"]]
        for chunk, want in zip(case["edge_chunks"], case["frames"]):
            edges = edges + [tuple(e) for e in chunk]
            got = []
            if len(edges) > 2:                                   # evalEdges, src/gpslib.py:1451-1462
                got = dec(None, edges)
                _, _, edges = navbits.logical_bits(edges)
            assert len(got) == len(want), case["name"]
            for g, w in zip(got, want):
                g = {k: (v.item() if hasattr(v, "item") else v) for k, v in g.items()}
                assert g == w, (case["name"], g, w)              # same keys, ints equal, floats bit-equal
        assert [int(b) for b in dec.bits] == case["kept_bits"], case["name"]
        assert [edges[0]] + [list(e) for e in edges[1:]] == case["kept_edges"], case["name"]


def test_encoder_round_trip_matches_the_quantised_ephemeris(gold):
    eph = gold["eph"]
    bits = navbits.encode_frames(777, 5, eph, first_id=1)
    assert bits.size == 1500 and set(np.unique(bits)) <= {0, 1}
    lsb = dict(Tgd=2.0 ** -31, af0=2.0 ** -31, af1=2.0 ** -43, Crs=2.0 ** -5, deltaN=2.0 ** -43 * navbits.GPS_PI,
               M0=2.0 ** -31 * navbits.GPS_PI, Cuc=2.0 ** -29, e=2.0 ** -33, Cus=2.0 ** -29, sqrtA=2.0 ** -19, Cic=2.0 ** -29,
               omegaBig=2.0 ** -31 * navbits.GPS_PI, Cis=2.0 ** -29, i0=2.0 ** -31 * navbits.GPS_PI, Crc=2.0 ** -5,
               omegaSmall=2.0 ** -31 * navbits.GPS_PI, omegaDot=2.0 ** -43 * navbits.GPS_PI, IDOT=2.0 ** -43 * navbits.GPS_PI)
    for k in range(5):
        status, f = navbits.decode_subframe(bits[300 * k:300 * (k + 1)])
        assert status == 0 and f["ID"] == k + 1 and f["tow"] == 777 + k
        for name, v in f.items():
            if name in lsb:
                assert abs(v - eph[name]) <= 0.5 * lsb[name] * (1 + 1e-12), name
            elif name in eph:
                assert v == eph[name], name
    inv = 1 - bits[:300]
    assert navbits.decode_subframe(inv)[1] == navbits.decode_subframe(bits[:300])[1]      # Costas ambiguity
    bad = bits[:300].copy()
    bad[200] ^= 1
    assert navbits.decode_subframe(bad)[0] == 3 and navbits.decode_subframe(bits[:299])[0] == 1


def test_random_ephemerides_round_trip_and_single_bit_errors_are_caught():
    """Property check: any quantised ephemeris survives encode -> decode; any single bit error in words 2..10 of a
    subframe is caught by the parity (word 1 is not parity-checked by the reference decoder, gpslib.py:379-405)."""
    rng = np.random.default_rng(2024)
    for trial in range(40):
        eph = dict(weekNum=int(rng.integers(0, 1024)), satAcc=int(rng.integers(0, 16)), satHealth=int(rng.integers(0, 64)),
                   Tgd=int(rng.integers(-128, 128)) * 2.0 ** -31, IODC=int(rng.integers(0, 1024)), Toc=int(rng.integers(0, 37800)) * 16,
                   af2=int(rng.integers(-128, 128)) * 2.0 ** -55, af1=int(rng.integers(-2 ** 15, 2 ** 15)) * 2.0 ** -43,
                   af0=int(rng.integers(-2 ** 21, 2 ** 21)) * 2.0 ** -31, IODE2=int(rng.integers(0, 256)),
                   Crs=int(rng.integers(-2 ** 15, 2 ** 15)) * 2.0 ** -5, deltaN=int(rng.integers(-2 ** 15, 2 ** 15)) * 2.0 ** -43 * navbits.GPS_PI,
                   M0=int(rng.integers(-2 ** 31, 2 ** 31)) * 2.0 ** -31 * navbits.GPS_PI, Cuc=int(rng.integers(-2 ** 15, 2 ** 15)) * 2.0 ** -29,
                   e=int(rng.integers(0, 2 ** 32)) * 2 ** -33, Cus=int(rng.integers(-2 ** 15, 2 ** 15)) * 2.0 ** -29,
                   sqrtA=int(rng.integers(0, 2 ** 32)) * 2.0 ** -19, Toe=int(rng.integers(0, 37800)) * 16,
                   Cic=int(rng.integers(-2 ** 15, 2 ** 15)) * 2.0 ** -29, omegaBig=int(rng.integers(-2 ** 31, 2 ** 31)) * 2.0 ** -31 * navbits.GPS_PI,
                   Cis=int(rng.integers(-2 ** 15, 2 ** 15)) * 2.0 ** -29, i0=int(rng.integers(-2 ** 31, 2 ** 31)) * 2.0 ** -31 * navbits.GPS_PI,
                   IODE3=int(rng.integers(0, 256)), Crc=int(rng.integers(-2 ** 15, 2 ** 15)) * 2.0 ** -5,
                   omegaSmall=int(rng.integers(-2 ** 31, 2 ** 31)) * 2.0 ** -31 * navbits.GPS_PI,
                   omegaDot=int(rng.integers(-2 ** 23, 2 ** 23)) * 2.0 ** -43 * navbits.GPS_PI,
                   IDOT=int(rng.integers(-2 ** 13, 2 ** 13)) * 2.0 ** -43 * navbits.GPS_PI)
        tow = int(rng.integers(0, 100800))
        for sf_id in (1, 2, 3, 4, 5):
            bits = np.asarray(navbits.encode_subframe(sf_id, tow, eph), dtype=np.int8)
            status, f = navbits.decode_subframe(bits)
            assert status == 0 and f["ID"] == sf_id and f["tow"] == tow
            assert all(f[k] == eph[k] for k in f if k in eph), (trial, sf_id, {k: (f[k], eph[k]) for k in f if k in eph and f[k] != eph[k]})
            pos = int(rng.integers(30, 300))
            bad = bits.copy()
            bad[pos] ^= 1
            assert navbits.decode_subframe(bad)[0] == 3, (trial, sf_id, pos)
