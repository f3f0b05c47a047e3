"""File / wire formats."""
from __future__ import annotations

import pickle

import numpy as np
import pytest

from gps_sdr_receiver_b200 import io as gio
from oracle import gps_oracle as orc


def test_read_streams_skips_and_drops_the_partial_tail(tmp_path):
    n_cyc, ngps = 8, 8 * 2048
    rng = np.random.default_rng(0)
    raw = rng.integers(0, 256, 2 * ngps * 5 + 1000, dtype=np.uint8)
    p = tmp_path / "rec.bin"
    raw.tofile(p)
    got = list(gio.read_streams(str(p), n_cyc, start_stream=2))
    assert len(got) == 3 and all(b.size == 2 * ngps for b in got)
    assert np.array_equal(got[0], raw[2 * ngps * 2:2 * ngps * 3])
    assert len(list(gio.read_streams(str(p), n_cyc, start_stream=0, max_streams=2))) == 2
    # the conversion is the reference reader's (bit-exact with the oracle's, itself pinned to the reference)
    assert np.array_equal(gio.raw_to_complex(got[1]), orc.raw_to_complex(got[1]))
    msg = gio.encode_message(0, [{"SAT": 5, "AMP": 1.0}], {5: [(3, 100.25)]})
    assert pickle.loads(msg) == (0, [{"SAT": 5, "AMP": 1.0}], {5: [(3, 100.25)]})
    with pytest.raises(ValueError):
        gio.send_udp(b"x" * 70000)


@pytest.mark.gpu
def test_file_receiver_messages_match_per_stream_processing(gpu, tmp_path):
    """A recording file with junk in front (START_STREAM) through FileReceiver: the messages carry the same frames
    and the same (streamNo, codePhase) pairs as feeding SatStream stream by stream like gpsrecv does."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from test_nav_e2e import N_CYC, _recording, _check_frames
    from gps_sdr_receiver_b200 import glob, navbits
    from gps_sdr_receiver_b200.tracking import SatStream
    sats, raw = _recording()
    ngps = N_CYC * 2048
    junk = np.full(2 * ngps * 3, 127, dtype=np.uint8)
    p = tmp_path / "rec.bin"
    np.concatenate([junk, raw]).tofile(p)
    rx = gio.FileReceiver(str(p), n_cyc=N_CYC, max_sat=4, start_stream=3, chunk_streams=16)
    msgs = list(rx.messages())
    assert {prn for _, prn, _, _ in rx.found} == {s.prn for s in sats}
    assert len(msgs) == (len(raw) // (2 * ngps)) // (1024 // N_CYC)             # one message per NO_SEC streams
    assert all(len(gio.encode_message(*m)) < gio.UDP_BUFSIZE for m in msgs)
    # per-stream reference flow with the same hand-over values
    glob.set_n_cyc(N_CYC)
    for z, prn, f, d in rx.found:
        sat = next(s for s in sats if s.prn == prn)
        ch = SatStream(prn, f, delay=d, frame_decoder=navbits.FrameDecoder())
        frames, coph = [], []
        for e in range(len(raw) // (2 * ngps)):
            smp = (e + 1) * ngps
            _, fl, cp, _ = ch.process(raw[2 * e * ngps:2 * (e + 1) * ngps], np.int64(smp))
            frames += fl
            if cp >= 0:
                coph.append((smp // ngps, float(cp)))
        ch.close()
        got_frames = [f for m in msgs for f in m[1] if f["SAT"] == prn]
        got_coph = [c for m in msgs for c in m[2].get(prn, [])]
        n_sent = len(got_coph)                                                   # code phases after the last message are not sent
        assert got_coph == coph[:n_sent] and n_sent >= len(coph) - 1024 // N_CYC
        assert got_frames == frames[:len(got_frames)] and len(got_frames) >= len(frames) - 1
        _check_frames([{k: v for k, v in f.items() if k not in ("SAT", "AMP", "CRM", "FRQ", "SWP")} for f in got_frames if "ID" in f],
                      sat, rolled=(prn == sats[1].prn))


@pytest.mark.gpu
def test_command_line_receiver_sends_the_messages_over_udp(gpu, tmp_path):
    import socket
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from test_nav_e2e import N_CYC, _recording
    sats, raw = _recording()
    p = tmp_path / "rec.bin"
    raw[:2 * N_CYC * 2048 * 160].tofile(p)                            # 5.1 s
    listener = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    listener.bind(("127.0.0.1", 0))
    listener.settimeout(5.0)
    port = listener.getsockname()[1]
    assert gio.main([str(p), "--port", str(port), "--n-cyc", str(N_CYC), "--max-sat", "4", "--quiet"]) == 0
    got = []
    for _ in range(160 // (1024 // N_CYC)):
        skipped, frame_lst, coph_lst = pickle.loads(listener.recv(gio.UDP_BUFSIZE))
        got.append((skipped, frame_lst, coph_lst))
    listener.close()
    assert len(got) == 5 and all(m[0] == 0 for m in got)
    assert {f["SAT"] for m in got for f in m[1]} == {s.prn for s in sats}
    assert all(len(v) > 0 for m in got for v in m[2].values())
