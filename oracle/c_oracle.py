"""ORACLE (test infrastructure) -- ctypes front-end of oracle/lap_oracle.c (level 2, batch capable).

Builds `oracle/_build/liblap_oracle.so` on demand with the recipe in oracle/Makefile.  Problem data
(track subset, vehicle constants) comes from the level-1 port's loaders so that the two oracles share
inputs but not arithmetic.  Same import restrictions as oracle/reference_port.py.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

from .reference_port import GRAV, OracleTrack

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblap_oracle.so")
_MAX_MAP = 16


class _Vehicle(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_map", C.c_int), ("mass", C.c_double), ("mu_g", C.c_double),
                ("f_max", C.c_double), ("f_max_sq", C.c_double),
                ("map_v", C.c_double * _MAX_MAP), ("map_f", C.c_double * _MAX_MAP),
                ("e0", C.c_double), ("cr2", C.c_double)]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("lap_oracle.c", "fitpack_port.c", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "--no-print-directory"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        _lib.lto_sweeps.restype = C.c_double
        _lib.lto_sweeps.argtypes = [C.POINTER(_Vehicle), dp, C.c_int, C.c_double, C.c_int, dp, dp, dp, dp]
        _lib.lto_eval.restype = C.c_int
        _lib.lto_eval.argtypes = [dp, dp, C.c_int, C.POINTER(_Vehicle), C.c_int, dp, C.c_long, dp,
                                  C.c_int, C.c_long, dp, dp, dp, dp, dp, dp]
        _lib.lto_path.restype = C.c_int
        _lib.lto_path.argtypes = [dp, dp, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
        _lib.lto_pairwise_sum.restype = C.c_double
        _lib.lto_pairwise_sum.argtypes = [dp, C.c_long]
        _lib.lto_set_use_pow.argtypes = [C.c_int]
        _lib.lto_set_sum_mode.argtypes = [C.c_int]
        _lib.lto_set_spline_mode.argtypes = [C.c_int]
        _lib.lto_pow15.restype = C.c_double
        _lib.lto_pow15.argtypes = [C.c_double]
        _lib.fpk_clocur.restype = C.c_int
        _lib.fpk_clocur.argtypes = [dp, dp, C.c_int, C.c_int, dp, dp]
        _lib.fpk_parcur_open.restype = C.c_int
        _lib.fpk_parcur_open.argtypes = [dp, dp, C.c_int, C.c_int, dp, dp]
        _lib.fpk_splder.restype = None
        _lib.fpk_splder.argtypes = [dp, C.c_int, dp, C.c_int, dp, C.c_int, dp, dp]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def vehicle_struct(veh):
    """Derived constants in the reference's own operation order (vehicle.py:30, vehicleMX5.py:28-33)."""
    v = _Vehicle()
    v.mass = veh.mass
    v.mu_g = veh.friction_coef * GRAV
    if veh.kind == "table":
        v.kind = 0
        f = veh.friction_coef * veh.mass * GRAV
        n = len(veh.map_v)
        assert n <= _MAX_MAP
        v.n_map = n
        for i in range(n):
            v.map_v[i] = veh.map_v[i]
            v.map_f[i] = veh.map_f[i]
    else:
        v.kind = 1
        D = (veh.D_f + veh.D_r) * 0.5
        Fn = veh.mass * GRAV
        f = 2.0 * D * Fn
        v.e0 = (veh.T * veh.C_m) - veh.Cr_0
        v.cr2 = veh.Cr_2
    v.f_max = f
    v.f_max_sq = f**2
    return v


class COracle:
    def __init__(self, track: OracleTrack, vehicle, mode="bayes", ns=None, use_pow=False,
                 device_sum_order=False, spline="tridiagonal"):
        self.track, self.vehicle, self.mode = track, vehicle, mode
        self.ns = math.ceil(track.length) if ns is None else ns
        if mode == "bayes":
            # effective polygon: all but the last every-3rd cone, closure implied
            self.left = np.ascontiguousarray(track.left_d[:, :-1])
            self.diff = np.ascontiguousarray(track.diffs_d[:, :-1])
        else:
            self.left = np.ascontiguousarray(track.left[:, :-1])
            self.diff = np.ascontiguousarray(track.diffs[:, :-1])
        self.N = self.left.shape[1]
        self.veh = vehicle_struct(vehicle)
        self.use_pow = use_pow
        # "tridiagonal": classical cyclic-tridiagonal periodic spline (the default CUDA K1a);
        # "fitpack": FITPACK's fpclos / splder arithmetic (oracle/fitpack_port.c), the reference's own bits
        self.spline_mode = {"tridiagonal": 0, "fitpack": 1}[spline]
        # True / "fused": summation order of the fused sweep kernel (the default device path);
        # "split": order of the separate backward kernel (LTK_SWEEP=split)
        self.sum_mode = {False: 0, None: 0, True: 2, "fused": 2, "split": 1}[device_sum_order]

    def _modes(self):
        L = lib()
        L.lto_set_use_pow(int(self.use_pow))
        L.lto_set_sum_mode(self.sum_mode)
        L.lto_set_spline_mode(self.spline_mode)
        return L

    def lap_times(self, alphas, threads=None):
        a = np.ascontiguousarray(alphas, dtype=np.float64).reshape(-1, self.N)
        lap = np.empty(a.shape[0])
        L = self._modes()
        rc = L.lto_eval(_p(self.left), _p(self.diff), self.N, C.byref(self.veh), self.ns, _p(a),
                        a.shape[0], _p(lap), threads or os.cpu_count(), -1, None, None, None, None,
                        None, None)
        assert rc == 0
        return lap

    def profile(self, alpha):
        a = np.ascontiguousarray(alpha, dtype=np.float64).reshape(1, self.N)
        n = self.ns - 1
        out = {k: np.empty(n) for k in ("k", "v_local", "v_acclim", "v_declim", "v")}
        lap, length = np.empty(1), np.empty(1)
        L = self._modes()
        rc = L.lto_eval(_p(self.left), _p(self.diff), self.N, C.byref(self.veh), self.ns, _p(a), 1,
                        _p(lap), 1, 0, _p(out["k"]), _p(out["v_local"]), _p(out["v_acclim"]),
                        _p(out["v_declim"]), _p(out["v"]), _p(length))
        assert rc == 0
        out["lap"], out["length"] = float(lap[0]), float(length[0])
        return out

    def sweeps(self, k, length, closed=True):
        k = np.ascontiguousarray(k, dtype=np.float64)
        n = k.size
        out = {key: np.empty(n) for key in ("v_local", "v_acclim", "v_declim", "v")}
        L = self._modes()
        out["lap"] = L.lto_sweeps(C.byref(self.veh), _p(k), n + 1, float(length), int(closed),
                                  _p(out["v_local"]), _p(out["v_acclim"]), _p(out["v_declim"]),
                                  _p(out["v"]))
        return out


def fitpack_spline(u, xy):
    """FITPACK periodic interpolating cubic spline (oracle/fitpack_port.c): u [m] parameters, xy [idim][m]
    points with xy[:, -1] == xy[:, 0].  Returns (t, [c_0, c_1, ...]) shaped like splprep's tck."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    xy = np.asarray(xy, dtype=np.float64)
    idim, m = xy.shape
    n = m + 6
    t, c = np.zeros(n), np.zeros(idim * n)
    pts = np.ascontiguousarray(xy.T.ravel())
    assert lib().fpk_clocur(_p(u), _p(pts), m, idim, _p(t), _p(c)) == n
    return t, [c[d * n:(d + 1) * n - 4].copy() for d in range(idim)]


def fitpack_spline_open(u, xy):
    """FITPACK open (not-a-knot) interpolating cubic spline: splprep(xy, u=u, k=3, s=0, per=0)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    xy = np.asarray(xy, dtype=np.float64)
    idim, m = xy.shape
    n = m + 4
    t, c = np.zeros(n), np.zeros(idim * n)
    pts = np.ascontiguousarray(xy.T.ravel())
    assert lib().fpk_parcur_open(_p(u), _p(pts), m, idim, _p(t), _p(c)) == n
    return t, [c[d * n:(d + 1) * n - 4].copy() for d in range(idim)]


def fitpack_splder(t, c, nu, x):
    """splev(x, (t, c, 3), der=nu) for one coordinate (oracle/fitpack_port.c splder)."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    cc = np.zeros(t.size)
    cc[:len(c)] = c
    x = np.ascontiguousarray(x, dtype=np.float64)
    y, wrk = np.zeros(x.size), np.zeros(t.size)
    lib().fpk_splder(_p(t), t.size, _p(cc), nu, _p(x), x.size, _p(y), _p(wrk))
    return y


def pow15(x):
    return lib().lto_pow15(float(x))


def path_derivatives(px, py, ns, spline="tridiagonal"):
    """Periodic spline of one closed polygon (unique points; closure implied)."""
    lib().lto_set_spline_mode({"tridiagonal": 0, "fitpack": 1}[spline])
    px = np.ascontiguousarray(px, dtype=np.float64)
    py = np.ascontiguousarray(py, dtype=np.float64)
    n = ns - 1
    out = {key: np.empty(n) for key in ("k", "dx", "dy", "ddx", "ddy")}
    length = np.empty(1)
    rc = lib().lto_path(_p(px), _p(py), px.size, ns, _p(length), _p(out["k"]), _p(out["dx"]),
                        _p(out["dy"]), _p(out["ddx"]), _p(out["ddy"]))
    assert rc == 0
    out["length"] = float(length[0])
    return out


def pairwise_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().lto_pairwise_sum(_p(a), a.size)
