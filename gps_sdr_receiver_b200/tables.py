"""Gold-code tables served by the CUDA library (host copies of the device tables).
`GPSCacode` is the drop-in for gpslib.GPSCacode (src/gpslib.py:70-77)."""
from __future__ import annotations

import numpy as np

from . import _capi


def chips(prn: int) -> np.ndarray:
    """cacodes.cacodes[prn] (src/cacodes.py): int8[1023] of +1/-1."""
    out = np.empty(1023, dtype=np.int8)
    _capi.check(_capi.lib().gr_get_chips(int(prn), out.ctypes.data))
    return out


def GPSCacode(satNo: int) -> np.ndarray:
    out = np.empty(2048, dtype=np.float64)
    _capi.check(_capi.lib().gr_get_cacode(int(satNo), out.ctypes.data))
    return out


def code_spectrum(prn: int) -> np.ndarray:
    """fft(GPSCacode(prn)) -- one FFT_CACODE entry (src/gpsrecv.py:574-577)."""
    out = np.empty(4096, dtype=np.float64)
    _capi.check(_capi.lib().gr_get_code_spectrum(int(prn), out.ctypes.data))
    return out.view(np.complex128)
