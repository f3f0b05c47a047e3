"""ctypes binding of include/gps_b200.h (the C ABI of libgpsb200.so).

The library is the product path; if it is missing or no CUDA device is present the
calls raise -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPSB200_LIB") or os.path.join(HERE, "libgpsb200.so")      # GPSB200_LIB: A/B runs against another build

GR_OK = 0
GR_IN_U8IQ, GR_IN_CF32 = 0, 1
GR_ACQ_ABS, GR_ACQ_POW = 0, 1
GR_MAX_PROMPT = 34

ACQ_CELL = np.dtype([("mx", "<i4"), ("peak", "<f4"), ("mean", "<f4"), ("std", "<f4"), ("z", "<f4"),
                     ("em1", "<f4"), ("ep1", "<f4"), ("second", "<f4")])
assert ACQ_CELL.itemsize == 32

EPOCH_OUT = np.dtype([
    ("prn", "<i4"), ("sweep", "<i4"), ("tracked", "<i4"), ("delay", "<i4"), ("corr_delay", "<i4"),
    ("locked", "<i4"), ("locked_in", "<i4"), ("report", "<i4"), ("rep_sweep", "<i4"), ("n_prompt", "<i4"),
    ("ms_time", "<i4"), ("n_prev", "<i4"), ("prompt_b1", "<i4"), ("freq_weak", "<i4"), ("edge0", "<i4"), ("edge_len", "<i4"),
    ("prompt_st0", "<i8"), ("edge_mask", "<u8"),
    ("code_phase", "<f8"), ("max_corr", "<f8"), ("corr_q", "<f8"), ("corr_l", "<f8"), ("freq", "<f8"),
    ("report_freq", "<f8"), ("phase", "<f8"), ("amplitude", "<f4"), ("std_dev", "<f4"), ("corr3", "<f4", (3,)),
    ("corr_mean", "<f4"), ("corr_std", "<f4"), ("erased", "<i4"),
    ("prompt", "<f4", (2 * GR_MAX_PROMPT,)), ("reserved", "<i4", (2,)),
], align=True)
assert EPOCH_OUT.itemsize == 448


ACQ_BEST = np.dtype([("prn", "<i4"), ("bin", "<i4"), ("cell", ACQ_CELL)])
assert ACQ_BEST.itemsize == 40


class SynthSat(C.Structure):
    _fields_ = [("prn", C.c_int32), ("bit_offset_ms", C.c_int32), ("bit_seed", C.c_uint32), ("amp", C.c_float),
                ("doppler", C.c_double), ("doppler_rate", C.c_double), ("delay", C.c_double), ("phi0", C.c_double)]


class SynthGeoSat(C.Structure):
    _fields_ = [("prn", C.c_int32), ("n_nodes", C.c_int32), ("n_bits", C.c_int32), ("amp", C.c_float),
                ("bit_t0_ms", C.c_int64), ("d_tau", C.c_void_p), ("d_bits", C.c_void_p)]


class TrackCfg(C.Structure):
    _fields_ = [("n_cyc", C.c_int32), ("corr_avg", C.c_int32), ("sweep_corr_avg", C.c_int32),
                ("it_sweep", C.c_int32), ("corr_min", C.c_float), ("min_freq", C.c_float),
                ("max_freq", C.c_float), ("step_freq", C.c_float), ("in_format", C.c_int32),
                ("max_channels", C.c_int32)]


class GrError(RuntimeError):
    pass


_lib = None

# every symbol include/gps_b200.h declares: (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "gr_version": (C.c_int, []),
    "gr_init": (C.c_int, [C.c_int]),
    "gr_shutdown": (C.c_int, []),
    "gr_last_error": (C.c_char_p, []),
    "gr_get_chips": (C.c_int, [C.c_int, _P]),
    "gr_get_cacode": (C.c_int, [C.c_int, _P]),
    "gr_get_code_spectrum": (C.c_int, [C.c_int, _P]),
    "gr_acq_classify_bins": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "gr_acq_plan_create": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "gr_acq_plan_destroy": (C.c_int, [_P]),
    "gr_acq_plan_form": (C.c_int, [_P]),
    "gr_acq_run_dev": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P, _P]),
    "gr_acq_run_host": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P]),
    "gr_acq_last_launches": (C.c_int, [_P]),
    "gr_acq_last_inverse_form": (C.c_int, [_P]),
    "gr_acq_search_dev": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P, _P]),
    "gr_acq_search_host": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P]),
    "gr_synth_iq_dev": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int, C.c_float, C.c_uint64, _P]),
    "gr_synth_geo_dev": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_float,
                                   C.c_uint64, _P]),
    "gr_debug_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "gr_track_default_cfg": (C.c_int, [C.POINTER(TrackCfg)]),
    "gr_track_bank_create": (C.c_int, [C.POINTER(TrackCfg), C.POINTER(_P)]),
    "gr_track_bank_destroy": (C.c_int, [_P]),
    "gr_track_add": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_int]),
    "gr_track_remove": (C.c_int, [_P, C.c_int]),
    "gr_track_request_sweep": (C.c_int, [_P, C.c_int]),
    "gr_track_process_dev": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int64, _P, _P]),
    "gr_track_process_host": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_int64, _P]),
    "gr_track_num_active": (C.c_int, [_P]),
    "gr_track_last_launches": (C.c_int, [_P]),
    "gr_track_bank_form": (C.c_int, [_P]),
    "gr_debug_fft2048": (C.c_int, [_P, _P, C.c_int, C.c_int]),
}


def lib() -> C.CDLL:
    """Load libgpsb200.so (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GrError(f"{LIB_PATH} is missing: build it with `python -m gps_sdr_receiver_b200._build` "
                          "(the CUDA library is the only implementation; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise GrError(f"libgpsb200 error {rc}: {lib().gr_last_error().decode(errors='replace')}")
    return rc


_initialised_device = None


def init(device: int = 0) -> None:
    global _initialised_device
    if _initialised_device != device:
        check(lib().gr_init(device))
        _initialised_device = device


def ptr(a) -> int:
    """Address of a numpy array or torch tensor (host or device)."""
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()
