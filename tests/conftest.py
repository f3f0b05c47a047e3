"""Shared fixtures.  `-m "not gpu"` runs everywhere; `-m gpu` needs a B200.

The oracle (oracle/gps_oracle.py) is imported only here and in the test modules: it
is the checker, never the product path."""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def built_lib():
    """libgpsb200.so, (re)built in-tree when nvcc is present and sources are newer."""
    from gps_sdr_receiver_b200 import _build, _capi
    try:
        _build.build()
    except RuntimeError:
        if not os.path.exists(_capi.LIB_PATH):
            raise
    return _capi.lib()


@pytest.fixture(scope="session")
def gpu(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a test marked gpu ran without a CUDA device")
    from gps_sdr_receiver_b200 import _capi
    _capi.init(0)
    return 0


class Scenario:
    """The seeded synthetic recording the golden fixtures were generated on
    (oracle/make_golden.py:scenario)."""

    def __init__(self, n_cyc: int):
        from gps_sdr_receiver_b200 import synth
        self.n_cyc = n_cyc
        self.gold = np.load(os.path.join(GOLD, f"traj_ncyc{n_cyc}.npz"))
        self.sats = synth.default_constellation(6, seed=5)
        self.n_epochs = int(self.gold["n_epochs"])
        self.raw = synth.make_iq(self.sats, n_cyc * self.n_epochs, noise_sigma=0.25, seed=11)
        assert _sha(self.raw) == str(self.gold["raw_sha"]), "synthetic generator drifted from the golden recording"
        self.ngps = n_cyc * 2048

    def block(self, e: int) -> np.ndarray:
        return self.raw[e * 2 * self.ngps:(e + 1) * 2 * self.ngps]


_SCEN = {}


@pytest.fixture(scope="session")
def scen32():
    if 32 not in _SCEN:
        _SCEN[32] = Scenario(32)
    return _SCEN[32]


@pytest.fixture(scope="session")
def scen8():
    if 8 not in _SCEN:
        _SCEN[8] = Scenario(8)
    return _SCEN[8]
