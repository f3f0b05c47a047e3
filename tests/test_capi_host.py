"""CPU-only checks of the C-ABI library: it loads, exports every symbol that
include/gps_b200.h declares, serves the code tables bit-exactly without a GPU, and
refuses to compute without one (no CPU fallback)."""
from __future__ import annotations

import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gps_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gr_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built_lib):
    from gps_sdr_receiver_b200 import _capi
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/gps_b200.h but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature in _capi.py"
    assert sorted(_capi.SIGNATURES) == names


def test_struct_layouts_match_header(built_lib):
    """numpy dtypes in _capi.py against sizeof() of the C structs (compiled with gcc)."""
    from gps_sdr_receiver_b200 import _capi
    src = ('#include <stdio.h>\n#include "gps_b200.h"\n'
           'int main(){printf("%zu %zu %zu\\n", sizeof(gr_acq_cell), sizeof(gr_epoch_out), sizeof(gr_track_cfg));return 0;}')
    exe = "/tmp/gr_sizeof"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    a, b, c = (int(v) for v in subprocess.run([exe], capture_output=True, check=True).stdout.split())
    import ctypes
    assert a == _capi.ACQ_CELL.itemsize
    assert b == _capi.EPOCH_OUT.itemsize
    assert c == ctypes.sizeof(_capi.TrackCfg)


def test_tables_bit_exact_without_gpu(built_lib):
    from gps_sdr_receiver_b200 import tables
    g = np.load(os.path.join(GOLD, "tables.npz"))
    chips = np.unpackbits(g["chips_packed"], axis=1)[:, :1023].astype(np.int8) * 2 - 1
    for p in range(1, 38):
        assert np.array_equal(tables.chips(p), chips[p - 1])
        assert np.array_equal(tables.GPSCacode(p), g["code_f32"][p - 1].astype(np.float64))
    for i, p in enumerate((1, 7, 19)):
        ref = g["spectrum_prn1_7_19"][i]
        got = tables.code_spectrum(p)
        assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max()
    with pytest.raises(Exception):
        tables.GPSCacode(0)
    with pytest.raises(Exception):
        tables.GPSCacode(38)


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gps_sdr_receiver_b200 import _capi
    rc = built_lib.gr_init(0)
    assert rc == -1
    assert b"no CPU fallback" in built_lib.gr_last_error()
    with pytest.raises(_capi.GrError):
        from gps_sdr_receiver_b200.acquisition import AcqPlan
        AcqPlan([2, 3], [0.0], 1)


def test_fft2048_index_mapping_host_emulation(tmp_path):
    """The three-stage CTA FFT (gr_fft2048.cuh) executed 'thread by thread' on the host."""
    exe = str(tmp_path / "fft_emul")
    subprocess.run(["g++", "-O2", "-I/usr/local/cuda/include", "-o", exe, os.path.join(ROOT, "tests", "host", "fft_emul.cpp")],
                   check=True)
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(2048) + 1j * rng.standard_normal(2048)).astype(np.complex64)
    y = np.frombuffer(subprocess.run([exe], input=x.tobytes(), capture_output=True, check=True).stdout, dtype=np.complex64)
    r = np.fft.fft(x.astype(np.complex128))
    assert np.abs(y - r).max() < 1e-6 * np.abs(r).max()


def test_two_instruction_time_base_equals_ieee_division_for_every_sample_index():
    """csrc/gr_common.cuh `tsec_of`: the kernels' reference-exact NCO forms need t[n] = fl32((n + 1) / fs) (gpsrecv.py:32-33,
    gpslib.py:1053-1054) per sample and form it as fma(k, y_hi, fl32(k * y_lo)) with 1 / fs = y_hi + y_lo.  Emulated here with
    exact products (k * y_hi has 48 significant bits, the sum fits a 64-bit mantissa, one rounding to float32) for EVERY
    k = 1 .. 2^23 against numpy's float32 division."""
    fs = np.float32(2048000.0)
    y_hi = np.array([889393775], dtype=np.uint32).view(np.float32)[0]
    y_lo = np.array([2832262496], dtype=np.uint32).view(np.float32)[0]
    assert y_hi == np.float32(1.0) / fs and y_lo == np.float32(1.0 / 2048000.0 - np.float64(y_hi))
    if np.finfo(np.longdouble).nmant < 63:
        pytest.skip("needs an 80-bit long double to emulate the fused multiply-add")
    bad = 0
    for lo in range(1, (1 << 23) + 1, 1 << 21):
        k = np.arange(lo, min(lo + (1 << 21), (1 << 23) + 1), dtype=np.float64)
        kf = k.astype(np.float32)
        c = kf * y_lo                                                   # float32 product, rounded
        t = (k.astype(np.longdouble) * np.longdouble(y_hi) + c.astype(np.longdouble)).astype(np.float32)
        bad += int((t != kf / fs).sum())
    assert bad == 0
