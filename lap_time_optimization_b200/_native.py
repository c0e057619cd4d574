"""ctypes binding of libltk.so (include/ltk.h).  There is no CPU fallback: if the CUDA library is
missing or fails to load, importing anything that computes raises `LtkUnavailable`."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LTK_LIB_PATH") or os.path.join(_HERE, "libltk.so")  # override: A/B builds of the library
MAX_ENGINE_MAP = 16

LTK_OK, LTK_E_ARG, LTK_E_CUDA, LTK_E_WORKSPACE, LTK_E_UNSUPPORTED = 0, -1, -2, -3, -4


class LtkUnavailable(RuntimeError):
    pass


class LtkError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"ltk error {code}: {message}")
        self.code = code


class LtkVehicle(C.Structure):
    """Mirror of `struct ltk_vehicle` (include/ltk.h)."""

    _fields_ = [("kind", C.c_int32), ("n_map", C.c_int32), ("mass", C.c_double), ("mu_g", C.c_double),
                ("f_max", C.c_double), ("f_max_sq", C.c_double),
                ("map_v", C.c_double * MAX_ENGINE_MAP), ("map_f", C.c_double * MAX_ENGINE_MAP),
                ("e0", C.c_double), ("cr2", C.c_double)]


# every symbol include/ltk.h declares: name -> (restype, argtypes)
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
SIGNATURES = {
    "ltk_create": (C.c_int, [C.POINTER(_vp), C.c_int, _dp, _dp, C.c_int, C.POINTER(LtkVehicle), C.c_int]),
    "ltk_destroy": (None, [_vp]),
    "ltk_last_error": (C.c_char_p, [_vp]),
    "ltk_set_ns": (C.c_int, [_vp, C.c_int]),
    "ltk_set_sweep_precision": (C.c_int, [_vp, C.c_int]),
    "ltk_set_spline_mode": (C.c_int, [_vp, C.c_int]),
    "ltk_spline_mode": (C.c_int, [_vp]),
    "ltk_trace_begin": (C.c_int, [_vp, C.c_int]),
    "ltk_trace_read": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                 C.POINTER(C.c_int)]),
    "ltk_set_sweep_split": (C.c_int, [_vp, C.c_int]),
    "ltk_workspace_bytes": (C.c_int, [_vp, C.c_int64, C.POINTER(C.c_size_t)]),
    "ltk_eval_alphas": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, C.c_size_t, _vp]),
    "ltk_eval_alphas_topk": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, C.c_size_t, C.c_int64, C.c_int, _vp, _vp, _vp]),
    "ltk_eval_alphas_host": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "ltk_eval_alphas_timed": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, C.c_size_t, _vp, C.POINTER(C.c_float)]),
    "ltk_eval_objectives": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ltk_random_uniform": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_double, C.c_double, _vp, _vp]),
    "ltk_topk_pairs": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp]),
    "ltk_topk_gathered": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "ltk_eval_controls": (C.c_int, [_vp, _vp, C.c_int, C.c_int64, _vp, _vp, C.c_size_t, _vp]),
    "ltk_profile": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ltk_topk": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, C.c_int, _vp, _vp, _vp]),
    "ltk_path_eval": (C.c_int, [C.c_int, _vp, _vp, C.c_int, _vp, C.c_int64] + [_vp] * 9),
    "ltk_path_eval_fitpack": (C.c_int, [C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, C.c_int64] + [_vp] * 11),
    "ltk_velocity_profile": (C.c_int, [C.c_int, C.POINTER(LtkVehicle), _vp, _vp, C.c_int64, C.c_double,
                                       _vp, _vp, _vp, _vp, _vp]),
    "ltk_version": (C.c_int, []),
    "ltk_launch_count": (C.c_int64, []),
}

_lib = None


def load():
    """Load libltk.so and attach signatures.  Raises LtkUnavailable when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LtkUnavailable(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise LtkUnavailable(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != LTK_OK:
        msg = load().ltk_last_error(ctx)
        raise LtkError(rc, msg.decode() if msg else "?")


def launch_count():
    return int(load().ltk_launch_count())
