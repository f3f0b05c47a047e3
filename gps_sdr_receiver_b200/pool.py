"""Drop-in for the multiprocessing worker pool of src/gpsrecv.py:300-417.

Same function names, arguments and return values as gpsrecv's
`initMultiProcPool / initPoolStreams / delPoolStreams / satCalc / closeMultiProcPool`,
so `gpsrecv.processData` (src/gpsrecv.py:445-548) runs unchanged on top.  Instead of one
OS process per satellite with the 512 kB stream pickled to each of them, all channels
live in one device-resident TrackBank and `satCalc` is ONE kernel launch."""
from __future__ import annotations

import numpy as np

from . import glob
from ._capi import GR_IN_CF32, GR_IN_U8IQ
from .tracking import SatStream, TrackBank


class ChannelPool:
    """What gpsrecv calls `pool`: the bank plus one host mirror (SatStream) per worker slot."""

    def __init__(self, poolNo: int, device: int = 0):
        self.poolNo = poolNo
        self.device = device
        self.bank: TrackBank | None = None
        self.streams: list[SatStream | None] = [None] * poolNo

    def _ensure_bank(self, fmt: int) -> TrackBank:
        if self.bank is None:
            self.bank = TrackBank(glob.N_CYC, self.poolNo, fmt, corr_avg=glob.CORR_AVG, sweep_corr_avg=glob.SWEEP_CORR_AVG,
                                  it_sweep=glob.IT_SWEEP, corr_min=glob.CORR_MIN, device=self.device)
        elif self.bank.in_format != fmt:
            raise TypeError("the pool must be fed one sample format (uint8 I/Q or complex64)")
        return self.bank

    def close(self):
        for s in self.streams:
            if s is not None:
                s.close()
        self.streams = [None] * self.poolNo
        if self.bank is not None:
            self.bank.close()
            self.bank = None


def initMultiProcPool(poolNo, in_format: int = GR_IN_CF32, device: int = 0):
    """gpsrecv.py:340-360.  Returns (pool, poolNo, poolWorker); poolWorker[w] = 0 if worker w is
    free, else the PRN it tracks."""
    pool = ChannelPool(poolNo, device)
    pool._ensure_bank(in_format)
    return pool, poolNo, [0] * poolNo


def closeMultiProcPool(pool):
    """gpsrecv.py:363-367."""
    pool.close()


def delPoolStreams(pool, poolNo, poolWorker, actSatSet, delSatSet):
    """gpsrecv.py:370-382."""
    for satNo in delSatSet:
        wno = poolWorker.index(satNo)
        if pool.streams[wno] is not None:
            pool.streams[wno].close()
            pool.streams[wno] = None
            poolWorker[wno] = 0
    return poolWorker, actSatSet - delSatSet


def initPoolStreams(pool, poolNo, poolWorker, actSatSet, newSatSet, foundSats):
    """gpsrecv.py:385-401: free workers take the new satellites with the (freq, delay) the
    cold-start search found for them."""
    if len(newSatSet) > 0:
        for wno, sno in enumerate(poolWorker):
            if sno == 0:
                newSat = newSatSet.pop()
                poolWorker[wno] = newSat
                _, _, freq, delay = list(filter(lambda e: e[1] == newSat, foundSats))[0]
                pool.streams[wno] = SatStream(newSat, freq, delay=delay, itSweep=glob.IT_SWEEP, corrMin=glob.CORR_MIN,
                                              corrAvg=glob.CORR_AVG, sweepCorrAvg=glob.SWEEP_CORR_AVG, bank=pool.bank)
                actSatSet.add(newSat)
                if len(newSatSet) == 0:
                    break
    return poolWorker, actSatSet


def satCalc(actSatSet, pool, poolWorker, data, smpTime):
    """gpsrecv.py:404-417: process one stream for every active satellite.  Returns the list of
    (swFq, satNo, frameData, coPh, cpQ) in actSatSet iteration order, like the reference."""
    if not actSatSet:
        return []
    a = np.asarray(data)
    fmt = GR_IN_U8IQ if a.dtype == np.uint8 else GR_IN_CF32
    if fmt == GR_IN_CF32 and a.dtype != np.complex64:
        a = a.astype(np.complex64)
    bank = pool._ensure_bank(fmt)
    recs = bank.process(a, int(smpTime), 1)[0]            # one launch for all channels
    by_slot = {slot: recs[i] for i, slot in enumerate(bank.slots)}
    resLst = []
    for sno in actSatSet:
        st = pool.streams[poolWorker.index(sno)]
        swFq, frameData, coPh, cpQ = st.absorb(by_slot[st._slot], smpTime)
        resLst.append((swFq, st.SAT_NO, frameData, coPh, cpQ))
    return resLst
