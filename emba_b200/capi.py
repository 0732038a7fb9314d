"""ctypes binding of libemba_b200.so (the C ABI declared in include/emba_b200.h).

This is plumbing: every call goes straight to the CUDA library. There is no CPU fallback -- if the shared
library is missing or no CUDA device is present the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libemba_b200.so")

OK = 0
STATE_CURRENT, STATE_CANDIDATE = 0, 1
COST_QUADRATIC, COST_CAUCHY, COST_HUBER = 0, 1, 2
ERRORS = {-1: "EMBA_E_ARG", -2: "EMBA_E_CUDA", -3: "EMBA_E_SUPPORT", -4: "EMBA_E_RANGE", -5: "EMBA_E_NCCL",
          -6: "EMBA_E_NUMERIC"}


class EmbaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("sensor_w", C.c_int32), ("sensor_h", C.c_int32), ("pano_w", C.c_int32), ("pano_h", C.c_int32),
                ("C_th", C.c_double), ("bearing_lut", C.POINTER(C.c_double)), ("device", C.c_int32)]


class LMSettings(C.Structure):
    _fields_ = [("max_num_iter", C.c_int32), ("tol_fun", C.c_double), ("num_times_tol_fun_sat", C.c_int32),
                ("use_cg", C.c_int32), ("cost_type", C.c_int32), ("eta", C.c_double),
                ("thres_valid_pixel", C.c_int32), ("damping_factor", C.c_double), ("alpha", C.c_double),
                ("first_time_window", C.c_int32)]


class LMLog(C.Structure):
    _fields_ = [("iter", C.c_int32), ("lambda_", C.c_double), ("cost_min", C.c_double), ("cost_new", C.c_double),
                ("accepted", C.c_int32), ("num_active_pixels", C.c_int64), ("num_measurements", C.c_int64),
                ("cg_iters", C.c_int32), ("cg_error", C.c_double), ("ms_form", C.c_double), ("ms_solve", C.c_double),
                ("ms_evaluate", C.c_double)]


LM_CALLBACK = C.CFUNCTYPE(C.c_int, C.POINTER(LMLog), C.c_void_p)


_dp = C.POINTER(C.c_double)
_H = C.c_void_p

# name -> (restype, argtypes); mirrors include/emba_b200.h one to one (tests/test_abi.py checks the header)
SIGNATURES = {
    "emba_create": (C.c_int, [C.POINTER(Config), C.POINTER(_H)]),
    "emba_destroy": (C.c_int, [_H]),
    "emba_last_error": (C.c_char_p, [_H]),
    "emba_set_events": (C.c_int, [_H, C.c_int64, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_uint8)]),
    "emba_num_pairs": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "emba_set_shard": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "emba_comm_unique_id": (C.c_int, [C.c_void_p]),
    "emba_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "emba_set_state": (C.c_int, [_H, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _dp, _dp, _dp]),
    "emba_get_state": (C.c_int, [_H, C.c_int32, _dp, _dp, _dp]),
    "emba_evaluate": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_double, C.c_double, _dp, _dp, C.POINTER(C.c_int64)]),
    "emba_get_evaluation": (C.c_int, [_H, C.c_int32, _dp, C.POINTER(C.c_int32)]),
    "emba_form_normal_eq": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_int64)]),
    "emba_set_map_path": (C.c_int, [_H, C.c_int32]),
    "emba_apply_l2_reg": (C.c_int, [_H, C.c_double]),
    "emba_get_normal_eq": (C.c_int, [_H, _dp, _dp, _dp, _dp, C.POINTER(C.c_int64), _dp]),
    "emba_a12_entries": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "emba_solve": (C.c_int, [_H, C.c_double, C.c_int32, C.c_int32, _dp, _dp, C.POINTER(C.c_int32), _dp]),
    "emba_make_candidate": (C.c_int, [_H, C.c_double, C.c_int32]),
    "emba_accept_candidate": (C.c_int, [_H]),
    "emba_solve_time_window": (C.c_int, [_H, C.POINTER(LMSettings), C.POINTER(LMLog), C.c_int32,
                                         C.POINTER(C.c_int32), _dp]),
    "emba_solve_time_window_cb": (C.c_int, [_H, C.POINTER(LMSettings), C.POINTER(LMLog), C.c_int32,
                                            C.POINTER(C.c_int32), _dp, LM_CALLBACK, C.c_void_p]),
    "emba_set_strict_range": (C.c_int, [_H, C.c_int32]),
    "emba_get_jacobian_rows": (C.c_int, [_H, _dp, C.c_int64, C.POINTER(C.c_int64)]),
    "emba_last_setup_ms": (C.c_int, [_H, _dp]),
    "emba_last_comm_ms": (C.c_int, [_H, _dp]),
    "emba_get_counters": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "emba_events_create": (C.c_int, [C.c_int32, C.c_int64, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16),
                                     C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.POINTER(C.c_void_p)]),
    "emba_events_destroy": (C.c_int, [C.c_void_p]),
    "emba_events_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "emba_events_sort_by_time": (C.c_int, [C.c_void_p]),
    "emba_events_subsample": (C.c_int, [C.c_void_p, C.c_int32]),
    "emba_events_window": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "emba_events_download": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_uint8)]),
    "emba_set_events_dev": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int64]),
    "emba_ext_create": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]),
    "emba_ext_destroy": (C.c_int, [C.c_void_p]),
    "emba_ext_num_pairs": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "emba_ext_set_state": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, _dp, _dp, _dp]),
    "emba_ext_get_state": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "emba_ext_evaluate": (C.c_int, [C.c_void_p, C.c_double, _dp, _dp, C.POINTER(C.c_int64)]),
    "emba_ext_form": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "emba_ext_get_rows": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int64)]),
    "emba_ext_get_normal_eq": (C.c_int, [C.c_void_p, _dp, _dp, _dp, C.POINTER(C.c_int32)]),
    "emba_ext_matvec": (C.c_int, [C.c_void_p, C.c_double, C.c_double, _dp, _dp]),
    "emba_ext_solve": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int32, C.c_double, _dp, C.POINTER(C.c_int32), _dp]),
    "emba_ext_apply": (C.c_int, [C.c_void_p, C.c_double]),
    "emba_ext_last_ms": (C.c_int, [C.c_void_p, _dp]),
    "emba_fit_control_poses": (C.c_int, [C.c_int32, C.c_int64, C.POINTER(C.c_int64), _dp, C.c_int64, C.c_int64,
                                         C.c_double, _dp, C.c_int32, C.POINTER(C.c_int32)]),
    "emba_poisson_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "emba_poisson_destroy": (C.c_int, [C.c_void_p]),
    "emba_poisson_reconstruct": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "emba_poisson_last_ms": (C.c_int, [C.c_void_p, _dp, C.POINTER(C.c_int64)]),
    "emba_reconstruct_map": (C.c_int, [_H, C.c_int32, _dp]),
    "emba_last_timings_ms": (C.c_int, [_H, _dp]),
    "emba_launch_count": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "emba_synchronize": (C.c_int, [_H]),
    "emba_version": (C.c_char_p, []),
}

_lib = None


def load():
    """Load libemba_b200.so and declare every entry point. Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not found: build it with `make -C emba_b200` "
                                    "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def ptr(a, ty=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))
