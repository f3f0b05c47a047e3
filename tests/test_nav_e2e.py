"""Valid LNAV subframes through signal synthesis -> tracking -> edge list -> nav-bit decoder
(SURVEY.md 8f N1 + N2 at bit level): the decoded frame dicts must carry the encoded ephemeris."""
from __future__ import annotations

import numpy as np
import pytest

from gps_sdr_receiver_b200 import navbits, synth
from oracle import gps_oracle as orc

EPH = dict(weekNum=345, satAcc=2, satHealth=0, Tgd=-1.1175870895385742e-08, IODC=0x2A5, Toc=230400, af2=0.0,
           af1=-3.637978807091713e-12, af0=-0.00021193176507949829, IODE2=0xA5, Crs=-88.65625, deltaN=4.3e-09,
           M0=-1.2345678, Cuc=-4.5e-06, e=0.0123456789, Cus=8.1e-06, sqrtA=5153.6789, Toe=230400, Cic=-1.1e-07,
           omegaBig=2.3456789, Cis=9.5e-08, i0=0.9765432, IODE3=0xA5, Crc=250.25, omegaSmall=-0.87654321,
           omegaDot=-8.1e-09, IDOT=2.5e-10)
N_CYC = 32
SECONDS = 14
FIRST_TOW = 4242


def _sats():
    bits = 2 * navbits.encode_frames(FIRST_TOW, 5, EPH, first_id=1).astype(np.int8) - 1      # 30 s, repeats
    return [synth.Sat(prn=9, doppler=1520.0, delay=700.4, amp=0.08, phi0=0.4, bit_offset_ms=11, bits=bits),
            synth.Sat(prn=23, doppler=-2210.0, delay=1650.7, amp=0.08, phi0=-1.0, bit_offset_ms=3, bits=np.roll(bits, 300))]


def _recording():
    sats = _sats()
    return sats, np.concatenate([synth.make_iq(sats, 1000, seed=3, start_sample=k * 1000 * 2048) for k in range(SECONDS)])


def _check_frames(frames, sat, rolled: bool):
    assert len(frames) >= 1, "no subframe decoded"
    want = navbits.encode_frames(FIRST_TOW, 5, EPH, first_id=1)
    ref = {k + 1: navbits.decode_subframe(want[300 * k:300 * (k + 1)])[1] for k in range(5)}
    for f in frames:
        st = int(f["ST"])
        g = {k: v for k, v in f.items() if k != "ST"}
        assert g == ref[f["ID"]], (g, ref[f["ID"]])                     # every field, floats bit-equal
        # the preamble's first bit starts at code start number bit_offset + 20 * 300 * m (+ 1 s shift for the rolled stream)
        sub_no = (f["ID"] - 1 + (1 if rolled else 0)) % 5
        first_ms = sat.bit_offset_ms + 6000 * sub_no
        # sample times count from SMP_TIME = NGPS at the first stream (gpsrecv.py:469-471), i.e. one epoch ahead of the index
        expect = sat.delay + 2048.0 * first_ms + N_CYC * 2048.0
        k = round((st - expect) / (2048.0 * 30000))                    # the 30-s message repeats
        assert abs(st - expect - k * 2048.0 * 30000) <= 2048.0, (st, expect)


def test_oracle_channel_plus_decoder_recovers_the_ephemeris():
    sats, raw = _recording()
    sat = sats[0]

    class DecodingChannel(orc.Channel):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.decoder, self.frames = navbits.FrameDecoder(), []

        def consume_edges(self):
            if len(self.edges) > 2:
                self.frames += self.decoder(None, list(self.edges))
            super().consume_edges()

    ch = DecodingChannel(sat.prn, 1500.0, delay=int(sat.delay) + 1, n_cyc=N_CYC)
    ngps = N_CYC * 2048
    for e in range(len(raw) // (2 * ngps)):
        ch.process(orc.raw_to_complex(raw[2 * e * ngps:2 * (e + 1) * ngps]), np.int64((e + 1) * ngps))
    assert ch.locked
    _check_frames(ch.frames, sat, rolled=False)


@pytest.mark.gpu
def test_gpu_satstream_with_frame_decoder_recovers_the_ephemeris(gpu):
    from gps_sdr_receiver_b200 import glob
    from gps_sdr_receiver_b200.tracking import SatStream
    glob.set_n_cyc(N_CYC)
    sats, raw = _recording()
    ngps = N_CYC * 2048
    for i, sat in enumerate(sats):
        ch = SatStream(sat.prn, 50.0 * round(sat.doppler / 50.0), delay=int(sat.delay) + 1, frame_decoder=navbits.FrameDecoder())
        frames = []
        for e in range(len(raw) // (2 * ngps)):
            _, fl, _, _ = ch.process(raw[2 * e * ngps:2 * (e + 1) * ngps], np.int64((e + 1) * ngps))
            frames += [f for f in fl if "ID" in f]
        assert ch.PHASE_LOCKED
        for f in frames:                                               # reportValues keys ride along (gpslib.py:1124-1131)
            assert f["SAT"] == sat.prn and "AMP" in f and "FRQ" in f
        _check_frames([{k: v for k, v in f.items() if k not in ("SAT", "AMP", "CRM", "FRQ", "SWP")} for f in frames], sat, rolled=(i == 1))
        ch.close()
