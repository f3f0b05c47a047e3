"""Receiver constants that drive the kernels -- the values of the reference's
src/gpsglob.py:38-42, 63-75, 119-131 (same names).  `set_n_cyc` replaces patching
gpsglob before import."""
import numpy as np

MAX_SAT = 11
IT_SWEEP = 40
IT_SWEEP_ALL = 10
CORR_AVG = 8
CORR_MIN = 8
SWEEP_CORR_AVG = 4
MIN_FREQ = -5000.0
MAX_FREQ = +5000.0
STEP_FREQ = 200
CODE_SAMPLES = 2048
SAMPLE_RATE = 1000 * CODE_SAMPLES
N_CYC = 32
NGPS = N_CYC * CODE_SAMPLES
MY_FLOAT = np.float32
MY_COMPLEX = np.complex64


def set_n_cyc(n_cyc: int) -> None:
    global N_CYC, NGPS
    if n_cyc not in (8, 16, 32):
        raise ValueError("N_CYC must be 8, 16 or 32 (gpsglob.py:122)")
    N_CYC = n_cyc
    NGPS = n_cyc * CODE_SAMPLES
