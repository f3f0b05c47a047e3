"""GPS LNAV bits on top of the tracking kernel's outputs.

Decoder -- what `gpslib.SatStream.evalEdges / logicalBits / evalGpsBits` and `gpslib.Subframe`
(src/gpslib.py:1451-1580, 96-419) do with the EDGES list that the tracker maintains: bit
edges -> 20-ms bits -> preamble search -> 300-bit subframes -> parity -> fields.  The result
dicts carry the same keys and the same float values as the reference's, so `gpseval` can
consume them unchanged.  `FrameDecoder` plugs into `tracking.SatStream(frame_decoder=...)`.

Encoder -- the inverse (IS-GPS-200 words with parity), which the reference does not have:
it produces the nav-bit streams of synthetic recordings whose subframes decode to known
ephemerides (tests, and the generator of SURVEY.md 8d).

Words are handled as Python ints: a 30-bit word is `d1 ... d24 D25 ... D30`, d1 the most
significant of its 24 data bits.  Parity (IS-GPS-200 table 20-XIV) is six masked XOR sums
over the data bits plus D29* / D30* of the previous word.
"""
from __future__ import annotations

import numpy as np

GPS_PI = 3.1415926535898                     # the ICD's pi (src/gpslib.py:16)
PREAMBLE_BITS = 0b10001011
_PREAMBLE_PM = np.array([1, -1, -1, -1, 1, -1, 1, 1], dtype=np.int8)     # +-1 form used by the correlator


def _mask(bits_1based) -> int:
    m = 0
    for b in bits_1based:
        m |= 1 << (24 - b)
    return m


# data-bit taps of D25..D30 (bit numbers 1..24), IS-GPS-200 20.3.5.2
_TAPS = (
    _mask((1, 2, 3, 5, 6, 10, 11, 12, 13, 14, 17, 18, 20, 23)),
    _mask((2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24)),
    _mask((1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22)),
    _mask((2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23)),
    _mask((1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24)),
    _mask((3, 5, 6, 8, 9, 10, 11, 13, 15, 19, 22, 23, 24)),
)
_PREV = (29, 30, 29, 30, 30, 29)             # which of D29* / D30* enters each parity bit


def word_parity(d24: int, p29: int, p30: int) -> int:
    """The six parity bits D25..D30 (as a 6-bit int) of source data bits d24."""
    out = 0
    for taps, prev in zip(_TAPS, _PREV):
        bit = (bin(d24 & taps).count("1") + (p29 if prev == 29 else p30)) & 1
        out = (out << 1) | bit
    return out


def encode_word(d24: int, p29: int, p30: int) -> int:
    """30 transmitted bits of one word: data bits complemented when D30* = 1, then parity."""
    par = word_parity(d24, p29, p30)
    tx = (d24 ^ 0xFFFFFF) if p30 else d24
    return (tx << 6) | par


def _solve_tail(d22: int, p29: int, p30: int) -> int:
    """Words 2 and 10 end with two non-information bits chosen so that D29 = D30 = 0: find them."""
    for t in range(4):
        d24 = (d22 << 2) | t
        if (word_parity(d24, p29, p30) & 3) == 0:
            return d24
    raise AssertionError("no tail bits give zero D29/D30")      # cannot happen: the 2 x 2 system is regular


def _u(value: int, nbits: int) -> int:
    return int(value) & ((1 << nbits) - 1)                    # two's complement for negative values


def _q(x: float, lsb: float) -> int:
    return int(round(x / lsb))


def encode_subframe(sf_id: int, tow: int, eph: dict | None = None, tlm_msg: int = 0) -> list[int]:
    """One 300-bit subframe as a list of 0/1 (transmission order).  `tow` is the 17-bit count of the
    NEXT subframe (what the HOW carries).  `eph`: field values in the units of the decoder's result
    dict (src/gpslib.py:316-371); subframes 4/5 carry alternating filler."""
    eph = eph or {}
    d = [0] * 10
    d[0] = (PREAMBLE_BITS << 16) | (_u(tlm_msg, 14) << 2)                       # TLM: preamble, message, 2 reserved
    how22 = (_u(tow, 17) << 5) | (0 << 3) | _u(sf_id, 3)                        # HOW: TOW, alert/AS = 0, id
    if sf_id == 1:
        iodc = _u(eph.get("IODC", 0), 10)
        d[2] = (_u(eph.get("weekNum", 0), 10) << 14) | (0 << 12) | (_u(eph.get("satAcc", 0), 4) << 8) | \
               (_u(eph.get("satHealth", 0), 6) << 2) | (iodc >> 8)
        d[3] = 0x555555 & 0xFFFFFF
        d[4] = 0x2AAAAA
        d[5] = 0x555555
        d[6] = (0xAAAA << 8) & 0xFFFF00 | _u(_q(eph.get("Tgd", 0.0), 2.0 ** -31), 8)
        d[7] = ((iodc & 0xFF) << 16) | _u(eph.get("Toc", 0) // 16, 16)
        d[8] = (_u(_q(eph.get("af2", 0.0), 2.0 ** -55), 8) << 16) | _u(_q(eph.get("af1", 0.0), 2.0 ** -43), 16)
        d9_22 = _u(_q(eph.get("af0", 0.0), 2.0 ** -31), 22)
    elif sf_id == 2:
        m0 = _u(_q(eph.get("M0", 0.0), 2.0 ** -31 * GPS_PI), 32)
        ecc = _u(_q(eph.get("e", 0.0), 2.0 ** -33), 32)
        sqa = _u(_q(eph.get("sqrtA", 0.0), 2.0 ** -19), 32)
        d[2] = (_u(eph.get("IODE2", 0), 8) << 16) | _u(_q(eph.get("Crs", 0.0), 2.0 ** -5), 16)
        d[3] = (_u(_q(eph.get("deltaN", 0.0), 2.0 ** -43 * GPS_PI), 16) << 8) | (m0 >> 24)
        d[4] = m0 & 0xFFFFFF
        d[5] = (_u(_q(eph.get("Cuc", 0.0), 2.0 ** -29), 16) << 8) | (ecc >> 24)
        d[6] = ecc & 0xFFFFFF
        d[7] = (_u(_q(eph.get("Cus", 0.0), 2.0 ** -29), 16) << 8) | (sqa >> 24)
        d[8] = sqa & 0xFFFFFF
        d9_22 = (_u(eph.get("Toe", 0) // 16, 16) << 6) | 0                       # fit interval flag, AODO = 0
    elif sf_id == 3:
        om0 = _u(_q(eph.get("omegaBig", 0.0), 2.0 ** -31 * GPS_PI), 32)
        i0 = _u(_q(eph.get("i0", 0.0), 2.0 ** -31 * GPS_PI), 32)
        w = _u(_q(eph.get("omegaSmall", 0.0), 2.0 ** -31 * GPS_PI), 32)
        d[2] = (_u(_q(eph.get("Cic", 0.0), 2.0 ** -29), 16) << 8) | (om0 >> 24)
        d[3] = om0 & 0xFFFFFF
        d[4] = (_u(_q(eph.get("Cis", 0.0), 2.0 ** -29), 16) << 8) | (i0 >> 24)
        d[5] = i0 & 0xFFFFFF
        d[6] = (_u(_q(eph.get("Crc", 0.0), 2.0 ** -5), 16) << 8) | (w >> 24)
        d[7] = w & 0xFFFFFF
        d[8] = _u(_q(eph.get("omegaDot", 0.0), 2.0 ** -43 * GPS_PI), 24)
        d9_22 = (_u(eph.get("IODE3", 0), 8) << 14) | _u(_q(eph.get("IDOT", 0.0), 2.0 ** -43 * GPS_PI), 14)
    else:
        for i in range(2, 9):
            d[i] = 0x555555 if i & 1 else 0xAAAAAA
        d9_22 = 0x155555
    words, p29, p30 = [], 0, 0                                  # word 1 follows a word 10, whose D29 = D30 = 0
    for i in range(10):
        if i == 1:
            d24 = _solve_tail(how22, p29, p30)
        elif i == 9:
            d24 = _solve_tail(d9_22, p29, p30)
        else:
            d24 = d[i]
        w30 = encode_word(d24, p29, p30)
        words.append(w30)
        p29, p30 = (w30 >> 1) & 1, w30 & 1
    return [(w >> (29 - b)) & 1 for w in words for b in range(30)]


def encode_frames(first_tow: int, n_subframes: int, eph: dict, first_id: int = 1) -> np.ndarray:
    """`n_subframes` consecutive subframes (ids cycling 1..5 from `first_id`) as an int8 array of 0/1.
    Subframe k carries HOW-TOW first_tow + k (the count of the subframe that follows it)."""
    bits = []
    for k in range(n_subframes):
        bits += encode_subframe((first_id - 1 + k) % 5 + 1, first_tow + k, eph)
    return np.asarray(bits, dtype=np.int8)


# ---- decoder ----------------------------------------------------------------------------------------

def _field(bits, signed=False) -> int:
    v = 0
    for b in bits:
        v = (v << 1) | int(b)
    if signed and int(bits[0]):
        v -= 1 << len(bits)
    return v


def decode_subframe(sub300) -> tuple[int, dict | None]:
    """(status, fields).  Status codes follow gpslib.Subframe (src/gpslib.py:97-108): 0 ok,
    1 length, 2 preamble, 3 parity, 4 id.  A subframe received inverted (Costas ambiguity) is
    accepted like the reference does; word 1 is not parity checked (src/gpslib.py:379-405)."""
    if len(sub300) != 300:
        return 1, None
    x = np.asarray(sub300, dtype=np.int64)
    pre = _field(x[:8])
    if pre == (PREAMBLE_BITS ^ 0xFF):
        x = 1 - x
    elif pre != PREAMBLE_BITS:
        return 2, None
    w = x.reshape(10, 30).copy()
    for i in range(1, 10):
        p29, p30 = int(w[i - 1, 28]), int(w[i - 1, 29])
        if p30:
            w[i, :24] = 1 - w[i, :24]
        if word_parity(_field(w[i, :24]), p29, p30) != _field(w[i, 24:]):
            return 3, None
    tow, sf_id = _field(w[1, :17]), _field(w[1, 19:22])
    if sf_id < 1 or sf_id > 5:
        return 4, None
    cat = lambda *parts: np.concatenate(parts)
    r = {"ID": sf_id, "tow": tow}
    if sf_id == 1:
        r.update(weekNum=_field(w[2, :10]), satAcc=_field(w[2, 12:16]), satHealth=_field(w[2, 16:22]),
                 Tgd=_field(w[6, 16:24], True) * 2 ** (-31), IODC=_field(cat(w[2, 22:24], w[7, :8])),
                 Toc=_field(w[7, 8:24]) * 16, af2=_field(w[8, 0:8], True) * 2.0 ** (-55),
                 af1=_field(w[8, 8:24], True) * 2.0 ** (-43), af0=_field(w[9, 0:22], True) * 2.0 ** (-31))
    elif sf_id == 2:
        r.update(Crs=_field(w[2, 8:24], True) * 2.0 ** (-5), deltaN=_field(w[3, 0:16], True) * 2.0 ** (-43) * GPS_PI,
                 M0=_field(cat(w[3, 16:24], w[4, 0:24]), True) * 2.0 ** (-31) * GPS_PI,
                 Cuc=_field(w[5, 0:16], True) * 2.0 ** (-29), IODE2=_field(w[2, 0:8]),
                 e=_field(cat(w[5, 16:24], w[6, 0:24])) * 2 ** (-33), Cus=_field(w[7, 0:16], True) * 2.0 ** (-29),
                 sqrtA=_field(cat(w[7, 16:24], w[8, 0:24])) * 2.0 ** (-19), Toe=_field(w[9, 0:16]) * 16)
    elif sf_id == 3:
        r.update(Cic=_field(w[2, 0:16], True) * 2.0 ** (-29),
                 omegaBig=_field(cat(w[2, 16:24], w[3, 0:24]), True) * 2.0 ** (-31) * GPS_PI,
                 Cis=_field(w[4, 0:16], True) * 2.0 ** (-29),
                 i0=_field(cat(w[4, 16:24], w[5, 0:24]), True) * 2.0 ** (-31) * GPS_PI, IODE3=_field(w[9, 0:8]),
                 Crc=_field(w[6, 0:16], True) * 2.0 ** (-5),
                 omegaSmall=_field(cat(w[6, 16:24], w[7, 0:24]), True) * 2.0 ** (-31) * GPS_PI,
                 omegaDot=_field(w[8, 0:24], True) * 2.0 ** (-43) * GPS_PI,
                 IDOT=_field(w[9, 8:22], True) * 2.0 ** (-43) * GPS_PI)
    return 0, r


def logical_bits(edges: list) -> tuple[np.ndarray, np.ndarray, list]:
    """EDGES = [first sign, (ms, sample time), ...] -> (+-1 bits, sample time of each run's first bit
    (0 elsewhere), trimmed EDGES).  A run of m bits lies between two edges (t2 - t1) ms apart with
    m = (t2 - t1) // 20, one more when the remainder exceeds 17 ms (src/gpslib.py:1465-1492)."""
    bits, st = [], []
    n = len(edges)
    if n <= 2:
        return np.zeros(0, np.int8), np.zeros(0, np.int64), edges
    sign = edges[0]
    t1, st1 = edges[1]
    for t2, st2 in edges[2:]:
        m, r = divmod(int(t2) - int(t1), 20)
        if r > 17:
            m += 1
        if m > 0:
            bits += [sign] * m
            st += [st1] + [0] * (m - 1)
        t1, st1, sign = t2, st2, -sign
    return np.asarray(bits, dtype=np.int8), np.asarray(st, dtype=np.int64), [sign, edges[-1]]


def find_subframes(bits_pm: np.ndarray, bits_st: np.ndarray) -> tuple[list[dict], int]:
    """Subframe dicts (with 'ST' = sample time of the preamble's first bit) found in a +-1 bit stream and
    the index from which the stream must be kept for the next call (src/gpslib.py:1504-1580)."""
    frames: list[dict] = []
    if len(bits_pm) < 300:
        return frames, 0
    corr = np.correlate(bits_pm.astype(np.int64), _PREAMBLE_PM.astype(np.int64), mode="same")
    locs = [int(i) - 4 for i in np.nonzero(np.abs(corr) == 8)[0]]
    start = 0
    if locs:
        gb = (bits_pm > 0).astype(np.int8)
        li, start, ok = 0, locs[0], True
        while ok and start + 300 < len(gb):
            status, fields = decode_subframe(gb[start:start + 300])
            if status == 0:
                fields["ST"] = bits_st[start]
                frames.append(fields)
                start += 300
            else:
                ok = False
                while not ok and li < len(locs) - 1:
                    li += 1
                    ok = locs[li] > start
                if ok:
                    start = locs[li]
    return frames, start


class FrameDecoder:
    """Per-channel bit memory (GPSBITS / GPSBITS_ST of the reference) + decode; use as
    `SatStream(..., frame_decoder=FrameDecoder())`: called with (stream, EDGES) once per second."""

    def __init__(self):
        self.bits = np.zeros(0, np.int8)
        self.st = np.zeros(0, np.int64)

    def reset(self):
        self.__init__()

    def __call__(self, stream, edges: list) -> list[dict]:
        b, s, _ = logical_bits(edges)
        self.bits = np.append(self.bits, b)
        self.st = np.append(self.st, s)
        frames, keep = find_subframes(self.bits, self.st)
        self.bits, self.st = self.bits[keep:], self.st[keep:]
        return frames
