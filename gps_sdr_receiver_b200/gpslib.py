"""`gpslib`-shaped facade: the names src/gpsrecv.py takes from the reference's gpslib on the
hot path (gpsrecv.py:316-320, 577), served by the CUDA library.

    import gps_sdr_receiver_b200.gpslib as gpslib     # in gpsrecv.py, instead of `import gpslib`

Everything off the hot path that gpseval/gpsui take from gpslib (SatOrbit, leastSquaresPos,
ecefToGeo, ...; gpseval.py:191,229,293,559-574) is resolved lazily from the reference's own
module when it is importable, so one import line serves both processes."""
from __future__ import annotations

from .tables import GPSCacode, chips, code_spectrum          # noqa: F401
from .tracking import SatStream, TrackBank                   # noqa: F401
from .navbits import FrameDecoder                            # noqa: F401  (evalEdges .. Subframe, gpslib.py:1451-1580)

import numpy as _np


def GPSCacodeRep(satNo, n_cyc, delay):
    """gpslib.GPSCacodeRep (src/gpslib.py:81-87): the 2048-sample code tiled n_cyc times, rolled."""
    return _np.roll(_np.tile(GPSCacode(satNo), n_cyc), delay)


def __getattr__(name):           # SatOrbit, Subframe, leastSquaresPos, ... : untouched reference code
    import importlib
    try:
        ref = importlib.import_module("gpslib")
    except ImportError as e:
        raise AttributeError(f"{name} is outside the B200 hot path and the reference's gpslib is not importable") from e
    if ref.__name__ == __name__:
        raise AttributeError(name)
    return getattr(ref, name)
