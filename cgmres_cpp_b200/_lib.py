"""ctypes binding of include/cgmres_b200.h -- loads the in-tree libcgmres_b200.so.

There is deliberately no fallback of any kind: if the shared library (built by
`__graft_entry__.build()` / `make -C cgmres_cpp_b200/csrc`) is missing, importing
the symbols raises, and creating a controller without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (CGMRES_B200_LIB: load an instrumented build of the same library instead, e.g. the -DCG_PIPE_TIMING one of
#  tools/pipe_wait_times.py; a debugging knob, not a fallback)
LIB_PATH = os.environ.get("CGMRES_B200_LIB") or os.path.join(_HERE, "libcgmres_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_h = C.c_void_p

# name -> (restype, argtypes); mirrors include/cgmres_b200.h one to one
SIGNATURES = {
    "cgmres_b200_last_error": (C.c_char_p, []),
    "cgmres_b200_device_count": (C.c_int, []),
    "cgmres_b200_model_dims": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "cgmres_b200_model_params": (C.c_int, [C.c_int, _dp]),
    "cgmres_b200_model_name": (C.c_char_p, [C.c_int]),
    "cgmres_b200_create": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.c_int, C.POINTER(_h)]),
    "cgmres_b200_destroy": (C.c_int, [_h]),
    "cgmres_b200_size": (C.c_int64, [_h]),
    "cgmres_b200_model": (C.c_int, [_h]),
    "cgmres_b200_mode": (C.c_int, [_h]),
    "cgmres_b200_set_stream": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_get_stream": (C.c_void_p, [_h]),
    "cgmres_b200_synchronize": (C.c_int, [_h]),
    "cgmres_b200_get_dtau": (C.c_double, [_h, C.c_double]),
    "cgmres_b200_set_ptau": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_set_ptau_repeat": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_init_u0": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_init_u0_newton": (C.c_int, [_h, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "cgmres_b200_control": (C.c_int, [_h, C.c_void_p, C.c_void_p]),
    "cgmres_b200_control_dev": (C.c_int, [_h, C.c_void_p, C.c_void_p]),
    "cgmres_b200_set_x": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_get_x": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_get_u": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_step_closed_loop": (C.c_int, [_h, C.c_int]),
    "cgmres_b200_step_closed_loop_log": (C.c_int, [_h, C.c_int, C.c_void_p, C.c_void_p]),
    "cgmres_b200_set_t": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_get_t": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_set_plant_integrator": (C.c_int, [_h, C.c_int]),
    "cgmres_b200_get_state": (C.c_int, [_h, _dp, C.c_void_p, C.c_void_p]),
    "cgmres_b200_set_state": (C.c_int, [_h, _dp, C.c_void_p, C.c_void_p]),
    "cgmres_b200_get_status": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_plant_step_host": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "cgmres_b200_portable_sincos": (None, [C.c_double, _dp, _dp]),
    "cgmres_b200_debug_phase_times": (C.c_int, [_h, C.c_void_p]),
    "cgmres_b200_launch_count": (C.c_int64, []),
    "cgmres_b200_measure_fp64_peak": (C.c_int, [C.c_int, C.c_int, _dp, _dp]),
    "cgmres_b200_measure_fp64_latency": (C.c_int, [C.c_int, C.c_int, _dp]),
}

_lib = None


class CgmresB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    """The loaded shared library with typed entry points; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CgmresB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C cgmres_cpp_b200/csrc` (there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().cgmres_b200_last_error()
        raise CgmresB200Error(f"cgmres_b200 error {rc}: {msg.decode() if msg else '?'}")
