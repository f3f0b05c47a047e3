"""File and wire formats around the hot path.

* recordings: the RTL-SDR byte stream the reference's recorder writes and `gpsrecv.streamData`
  reads (src/gpsrecv.py:153-186): one uint16 little-endian word per sample = (I byte, Q byte);
  a "stream" is NGPS = N_CYC * 2048 samples; START_STREAM whole streams are skipped, a
  trailing partial stream ends the file.
* messages: what gpsrecv sends to gpseval once per second (src/gpsrecv.py:496-519):
  `pickle.dumps((skippedData, frameLst, coPhLst))` over UDP, frameLst = the frame / report dicts of
  all channels, coPhLst = {satNo: [(streamNo, codePhase), ...]} accumulated since the last message.

`FileReceiver` is the headless equivalent of `gpsrecv.processData` for a file: a fine cold-start search
on the first streams, then all channels of every chunk of streams in ONE kernel launch (pinned
double-buffered upload inside `gr_track_process_host`), yielding those messages.
"""
from __future__ import annotations

import pickle
import socket

import numpy as np

from . import glob
from .navbits import FrameDecoder
from .tracking import SatStream, TrackBank

UDP_PORT = 61431                 # gpsglob.py:82
UDP_BUFSIZE = 65504              # gpsglob.py:85


def read_streams(path: str, n_cyc: int | None = None, start_stream: int = 0, max_streams: int | None = None):
    """Yield uint8[2 * NGPS] blocks (I,Q bytes) of a recording, like streamData does."""
    ngps = (glob.N_CYC if n_cyc is None else n_cyc) * glob.CODE_SAMPLES
    with open(path, "rb") as f:
        f.seek(2 * ngps * start_stream)
        k = 0
        while max_streams is None or k < max_streams:
            block = np.fromfile(f, dtype=np.uint8, count=2 * ngps)
            if block.size != 2 * ngps:
                return
            yield block
            k += 1


def raw_to_complex(block: np.ndarray) -> np.ndarray:
    """The reference reader's conversion (gpsrecv.py:168-173): complex64 in [-1, 1]."""
    im, re = np.divmod(np.ascontiguousarray(block).view(np.uint16), 256)
    return np.asarray(re + 1j * im, dtype=np.complex64) / 127.5 - (1 + 1j)


def encode_message(skipped: int, frame_lst: list, coph_lst: dict) -> bytes:
    return pickle.dumps((skipped, frame_lst, coph_lst))


def send_udp(payload: bytes, ip: str = "127.0.0.1", port: int = UDP_PORT, sock: socket.socket | None = None) -> None:
    if len(payload) > UDP_BUFSIZE:
        raise ValueError(f"message of {len(payload)} bytes exceeds gpseval's UDP buffer ({UDP_BUFSIZE})")
    s = sock or socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    try:
        s.sendto(payload, (ip, port))
    finally:
        if sock is None:
            s.close()


class FileReceiver:
    def __init__(self, path: str, n_cyc: int = 32, max_sat: int = glob.MAX_SAT, start_stream: int = 0, chunk_streams: int = 32,
                 z_min: float = 18.0, device: int = 0):
        self.path, self.n_cyc, self.max_sat, self.start, self.chunk = path, int(n_cyc), int(max_sat), int(start_stream), int(chunk_streams)
        self.z_min, self.device = float(z_min), device
        self.ngps = self.n_cyc * glob.CODE_SAMPLES
        self.found: list[tuple[float, int, float, int]] = []          # (z, prn, freq, delay) like gpsrecv.foundSats

    ACQ_TCOH_MS, ACQ_NNONCOH = 10, 2                                   # fine cold-start search: 20 ms of samples

    def _cold_start(self, lead: np.ndarray):
        """`lead`: the first streams of the recording, at least ACQ_TCOH_MS * ACQ_NNONCOH ms of samples."""
        from .acquisition import AcqPlan, GR_ACQ_POW
        from .batch import select_sats
        prns = list(range(1, 33))
        bins = [glob.MIN_FREQ + 50.0 * b for b in range(int((glob.MAX_FREQ - glob.MIN_FREQ) / 50.0) + 1)]
        plan = AcqPlan(prns, bins, self.ACQ_TCOH_MS, self.ACQ_NNONCOH, GR_ACQ_POW, device=self.device)
        try:
            best = plan.search(lead[:2 * plan.rec_samples])[0]
        finally:
            plan.close()
        self.found = sorted(((float(best[i]["cell"]["z"]), prns[i], bins[int(best[i]["bin"])], int(best[i]["cell"]["mx"]))
                             for i in select_sats(best, self.z_min, self.max_sat)), reverse=True)

    def messages(self):
        """Yield (skippedData, frameLst, coPhLst) tuples, one per second of recording (gpsrecv.py:496-519)."""
        glob.set_n_cyc(self.n_cyc)
        blocks = read_streams(self.path, self.n_cyc, self.start)
        # the search window is 20 ms: one stream at N_CYC = 32, two at 16, three at 8 -- they are tracked afterwards like the rest
        need = -(-self.ACQ_TCOH_MS * self.ACQ_NNONCOH * glob.CODE_SAMPLES // self.ngps)
        lead = []
        for b in blocks:
            lead.append(b)
            if len(lead) == need:
                break
        if len(lead) < need:
            return                                                      # shorter than the search window: nothing to report
        self._cold_start(np.concatenate(lead))
        if not self.found:
            return
        bank = TrackBank(self.n_cyc, len(self.found), device=self.device)
        streams = [SatStream(prn, f, delay=d, bank=bank, frame_decoder=FrameDecoder()) for _, prn, f, d in self.found]
        coph: dict = {}
        smp = self.ngps                                                # SMP_TIME of the first stream (gpsrecv.py:469-471)
        pending = lead
        try:
            while pending:
                while len(pending) < self.chunk:
                    b = next(blocks, None)
                    if b is None:
                        break
                    pending.append(b)
                raw = np.concatenate(pending)
                recs = bank.process(raw, smp, n_epochs=len(pending))
                for e in range(len(pending)):
                    frame_lst, stream_no = [], smp // self.ngps
                    for c, st in enumerate(streams):
                        _, f_lst, co_ph, _ = st.absorb(recs[e, c], smp)
                        frame_lst += f_lst
                        if co_ph >= 0:
                            coph.setdefault(st.SAT_NO, []).append((int(stream_no), float(co_ph)))
                    if frame_lst:
                        yield 0, frame_lst, coph
                        coph = {}
                    smp += self.ngps
                nxt = next(blocks, None)
                pending = [] if nxt is None else [nxt]
        finally:
            for st in streams:
                st.close()
            bank.close()


def main(argv=None) -> int:
    """`python -m gps_sdr_receiver_b200.io recording.bin [--ip 127.0.0.1] [--port 61431] [--start-stream N] [--n-cyc 32]`:
    the headless replacement of `gpsrecv.py` for a recorded file -- the reference's `gpseval.py` (and its GUI) can
    listen on the UDP port unchanged."""
    import argparse
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("recording")
    ap.add_argument("--ip", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=UDP_PORT)
    ap.add_argument("--start-stream", type=int, default=0)
    ap.add_argument("--n-cyc", type=int, default=32)
    ap.add_argument("--max-sat", type=int, default=glob.MAX_SAT)
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args(argv)
    rx = FileReceiver(a.recording, n_cyc=a.n_cyc, max_sat=a.max_sat, start_stream=a.start_stream)
    sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    n = 0
    try:
        for skipped, frame_lst, coph_lst in rx.messages():
            send_udp(encode_message(skipped, frame_lst, coph_lst), a.ip, a.port, sock)
            n += 1
            if not a.quiet:
                ids = [(f["SAT"], f["ID"], f["tow"]) for f in frame_lst if "ID" in f]
                print(f"message {n}: {len(frame_lst)} frame dicts, subframes {ids}, code phases of {sorted(coph_lst)}")
    finally:
        sock.close()
    if not a.quiet:
        print(f"satellites tracked: {[(prn, f) for _, prn, f, _ in rx.found]}; {n} messages sent to {a.ip}:{a.port}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
