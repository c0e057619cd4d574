"""ORACLE (test infrastructure, not product code) -- level 1: scalar CPU restatement of the reference.

A from-scratch, one-candidate-at-a-time restatement of the reference's lap-time path in plain
Python / numpy / SciPy-FITPACK, i.e. with the SAME third-party arithmetic the reference uses
(`scipy.interpolate.splprep/splev`, `np.interp`, libm `pow`/`sqrt`).  Every function cites the
reference file:line it follows (paths relative to /root/reference).

Parity status: PINNED.  `tests/golden/*.npz` were produced by importing the unmodified reference in
the build container (`tools/make_golden.py`); `tests/test_oracle.py` checks this port against them
bit-for-bit (lap, curvature, v_local, v_acclim, v_declim, v).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this module.  The product (`lap_time_optimization_b200/`) never does.
"""
from __future__ import annotations

import json
import math
import re

import numpy as np
from scipy.interpolate import splev, splprep

GRAV = 9.81  # velocity.py:4, vehicle.py:5, vehicleMX5.py:6


# --------------------------------------------------------------------------------------------
# problem data
# --------------------------------------------------------------------------------------------
class OracleVehicle:
    """TBR18-style point mass with a tabulated engine map (vehicle.py:11-35)."""

    kind = "table"

    def __init__(self, path):
        d = json.load(open(path))
        self.name = d["name"]
        self.mass = d["mass"]
        self.friction_coef = d["frictionCoefficient"]
        self.map_v = d["engineMap"]["v"]
        self.map_f = d["engineMap"]["f"]

    def engine_force(self, v):  # vehicle.py:25-27
        return np.interp(v, self.map_v, self.map_f)

    def traction(self, v, k):  # vehicle.py:29-35
        f = self.friction_coef * self.mass * GRAV
        f_lat = self.mass * v**2 * k
        if f <= f_lat:
            return 0
        return math.sqrt(f**2 - f_lat**2)


class OracleVehicleMX5:
    """MX-5 reduced to the same two callbacks (vehicleMX5.py:19-37, :46-79)."""

    kind = "mx5"

    def __init__(self, path):
        txt = open(path).read()
        txt = re.sub(r"//.*", "", txt)
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.DOTALL)
        d = json.loads(txt)
        self.name = d["name"]
        self.mass = d["mass"]
        self.D_f = d["frontTire"]["D_f"]
        self.D_r = d["rearTire"]["D_r"]
        self.C_m = d["control"]["C_m"]
        self.T = d["control"]["T"]
        self.Cr_0 = d["Cr_0"]
        self.Cr_2 = d["Cr_2"]
        self.friction_coef = d["control"]["lambda"]  # vehicleMX5.py:77

    def engine_force(self, v):  # vehicleMX5.py:19-21
        return (self.T * self.C_m) - self.Cr_0 - (self.Cr_2 * (v**2))

    def traction(self, v, k, lam=2.0):  # vehicleMX5.py:23-37
        D = (self.D_f + self.D_r) * 0.5
        Fn = self.mass * GRAV
        F_max = lam * D * Fn
        F_lat = self.mass * v * v * k
        if F_max <= F_lat:
            return 0
        return math.sqrt(F_max**2 - F_lat**2)


def load_vehicle(path):
    """MX5 files have a `control` block, TBR18 files an `engineMap` (the reference picks the class by
    comparing the path string, __main__.py:100; we look at the content)."""
    txt = re.sub(r"//.*", "", open(path).read())
    return OracleVehicleMX5(path) if '"control"' in txt else OracleVehicle(path)


def _chord_knots(points):
    """Cumulative chord length (path.py:11-14)."""
    seg = np.linalg.norm(np.diff(points, axis=1), axis=0)
    return np.append(0, np.cumsum(seg))


class OracleTrack:
    """Cone boundaries, shrunk to a width fraction, plus the every-3rd-cone subset (track.py:11-49)."""

    def __init__(self, json_path, track_width):
        d = json.load(open(json_path))
        self.name = d["name"]
        left = np.array([d["left"]["x"], d["left"]["y"]])
        right = np.array([d["right"]["x"], d["right"]["y"]])
        # track.py:17-21 clamps the fraction and stores its complement
        w = min(max(track_width, 0.001), 1.0)
        inv = 1.0 - w
        # track.py:96-118: each side moves inwards by inv/2 of the cone-to-cone vector
        new_left = np.zeros_like(left)
        new_right = np.zeros_like(right)
        for i in range(left.shape[1]):
            new_left[:, i] = left[:, i] + inv * (right[:, i] - left[:, i]) / 2
            new_right[:, i] = right[:, i] + inv * (left[:, i] - right[:, i]) / 2
        self.left, self.right = new_left, new_right
        # utils.py:17-22
        self.closed = bool(all(self.left[:, 0] == self.left[:, -1])
                           and all(self.right[:, 0] == self.right[:, -1]))
        self.size = self.left.shape[1] - int(self.closed)  # track.py:24
        self.diffs = self.right - self.left  # track.py:25
        mid = self.control_points(np.full(self.size, 0.5))
        self.length = _chord_knots(mid)[-1]  # track.py:27-28
        sel = np.arange(0, mid.shape[1], 3)  # track.py:40
        self.left_d = self.left[:, sel]
        self.diffs_d = self.diffs[:, sel]

    def control_points(self, alphas):  # track.py:82-87
        alphas = np.asarray(alphas, dtype=float)
        if self.closed:
            alphas = np.append(alphas, alphas[0])
        return self.left + alphas * self.diffs

    def control_points_bayesian(self, alphas):  # track.py:89-94
        alphas = np.asarray(alphas, dtype=float)
        if self.closed:
            alphas = np.append(alphas, alphas[0])
        return self.left_d + alphas * self.diffs_d


# --------------------------------------------------------------------------------------------
# geometry: FITPACK periodic spline, sampling, curvature
# --------------------------------------------------------------------------------------------
class OraclePath:
    """path.py:17-61.  NOTE: with per=1 scipy's splprep overwrites controls[:, -1] with
    controls[:, 0] in place; the reference relies on that side effect."""

    def __init__(self, controls, closed):
        self.controls = controls
        self.closed = closed
        self.dists = _chord_knots(controls)
        self.spline, _ = splprep(controls, u=self.dists, k=3, s=0, per=closed)
        self.length = self.dists[-1]

    def curvature(self, u):  # path.py:36-61
        dx, dy = splev(u, self.spline, der=1)
        ddx, ddy = splev(u, self.spline, der=2)
        return np.abs((dx * ddy - dy * ddx) / (dx**2 + dy**2) ** (3 / 2))

    def gamma2(self, u):  # path.py:63-77
        dx, dy = splev(u, self.spline, der=1)
        ddx, ddy = splev(u, self.spline, der=2)
        k = (dx * ddy - dy * ddx) / (dx**2 + dy**2) ** (3 / 2)
        return np.sum(k**2)

    def position(self, u):  # path.py:29-34
        x, y = splev(u, self.spline)
        return np.array([x, y])


# --------------------------------------------------------------------------------------------
# dynamics: three-pass velocity profile
# --------------------------------------------------------------------------------------------
def velocity_profile(vehicle, s, k, s_max):
    """velocity.py:14-76 restated with explicit sample indices instead of roll/flip.

    Returns (v_local, v_acclim, v_declim, v).  `s`, `k` exclude the closing sample; `s_max` is the
    period for a closed path or None for an open one.
    """
    n = s.size
    v_local = np.sqrt(vehicle.friction_coef * GRAV / k)  # velocity.py:28-29
    p = int(np.argmin(v_local))  # velocity.py:34, :58 (first minimum)
    mass = vehicle.mass

    # forward pass, velocity.py:31-53: visit p, p+1, ..., p+n-1 (mod n); sample 0 is the wrap step
    va = v_local.copy()
    for i in range(n):
        q = (p + i) % n
        prev = (q - 1) % n
        wrap = q == 0
        if wrap and s_max is None:
            continue
        if va[q] > va[prev]:
            tr = vehicle.traction(va[prev], k[prev])
            force = min(vehicle.engine_force(va[prev]), tr)
            accel = force / mass
            ds = s_max - s[prev] if wrap else s[q] - s[prev]
            vlim = math.sqrt(va[prev] ** 2 + 2 * accel * ds)
            va[q] = min(va[q], vlim)

    # backward pass, velocity.py:55-76: visit p, p-1, ..., p-n+1 (mod n); the wrap step is the one
    # whose predecessor-in-time (next in space) is sample 0, i.e. q == n-1
    vd = v_local.copy()
    for i in range(n):
        q = (p - i) % n
        nxt = (q + 1) % n
        wrap = q == n - 1
        if wrap and s_max is None:
            continue
        if vd[q] > vd[nxt]:
            tr = vehicle.traction(vd[nxt], k[nxt])
            decel = tr / mass
            ds = s_max - s[q] if wrap else s[nxt] - s[q]
            vlim = math.sqrt(vd[nxt] ** 2 + 2 * decel * ds)
            vd[q] = min(vd[q], vlim)

    return v_local, va, vd, np.minimum(va, vd)  # velocity.py:26


# --------------------------------------------------------------------------------------------
# the evaluator (the reference's Trajectory / TrajectoryBayesianNonlinear call surface)
# --------------------------------------------------------------------------------------------
class OracleEvaluator:
    """alphas -> lap time, one candidate at a time.

    mode "bayes": trajectory_bayesian_nonlinear.py:58-80 (`calcMinTime(updateAlphas(a))`, every-3rd-cone
                  control points, including the in-place closure quirk);
    mode "full" : trajectory.py:40-58 (`update`, `update_velocity`, `lap_time`, one alpha per cone).
    """

    def __init__(self, track: OracleTrack, vehicle, mode="bayes", ns=None):
        self.track = track
        self.vehicle = vehicle
        self.mode = mode
        self.ns = math.ceil(track.length) if ns is None else ns  # trajectory.py:35, tbn.py:31

    def n_alpha(self):
        return self.track.left_d.shape[1] - 1 if self.mode == "bayes" else self.track.size

    def controls(self, alphas):
        if self.mode == "bayes":
            c = self.track.control_points_bayesian(alphas)
            OraclePath(c, self.track.closed)  # tbn.py:59: the discarded Path whose splprep closes c in place
            return c
        return self.track.control_points(alphas)

    def profile(self, alphas):
        """Everything the reference computes for one candidate, as a dict."""
        c = self.controls(alphas)
        path = OraclePath(c, self.track.closed)  # tbn.py:69 / trajectory.py:43
        s = np.linspace(0, path.length, self.ns)  # tbn.py:71 / trajectory.py:45
        s_max = path.length if self.track.closed else None
        k = path.curvature(s[:-1])  # tbn.py:76
        v_local, va, vd, v = velocity_profile(self.vehicle, s[:-1], k, s_max)  # tbn.py:77
        lap = np.sum(np.diff(s) / v)  # tbn.py:51-54
        return dict(controls=c, length=path.length, s=s, k=k, v_local=v_local, v_acclim=va,
                    v_declim=vd, v=v, lap=lap, path=path)

    def lap_time(self, alphas):
        return float(self.profile(alphas)["lap"])


def top_k(laps, k):
    """tbn.py:253-257: `sorted(zip(taus, alphas), key=tau)[0:k]` -- stable ascending."""
    order = sorted(range(len(laps)), key=lambda i: laps[i])[:k]
    return np.asarray(order, dtype=np.int64), np.asarray([laps[i] for i in order], dtype=np.float64)


# ---- multi-process batch helper (CPU baseline; fork pool, one evaluator per worker) -------------
_WORKER = None


def _init_worker(track_json, width, vehicle_json, mode, ns):
    global _WORKER
    _WORKER = OracleEvaluator(OracleTrack(track_json, width), load_vehicle(vehicle_json), mode, ns)


def _eval_rows(rows):
    return [_WORKER.lap_time(a) for a in rows]


class LapPool:
    """A persistent fork pool of `processes` workers, one reference-equivalent evaluator each."""

    def __init__(self, track_json, width, vehicle_json, mode="bayes", ns=None, processes=None):
        import multiprocessing as mp
        import os

        self.processes = processes or os.cpu_count() or 1
        self.pool = mp.get_context("fork").Pool(self.processes, initializer=_init_worker,
                                                 initargs=(track_json, width, vehicle_json, mode, ns))

    def lap_times(self, alphas):
        alphas = np.asarray(alphas, dtype=np.float64)
        chunks = np.array_split(alphas, min(len(alphas), self.processes * 8))
        out = self.pool.map(_eval_rows, chunks)
        return np.asarray([x for part in out for x in part])

    def close(self):
        self.pool.close()
        self.pool.join()


def lap_times_pool(track_json, width, vehicle_json, alphas, mode="bayes", ns=None, processes=None):
    """Score the rows of `alphas` with `processes` worker processes; returns laps[B]."""
    import os

    processes = processes or os.cpu_count()
    alphas = np.asarray(alphas, dtype=np.float64)
    if processes <= 1:
        _init_worker(track_json, width, vehicle_json, mode, ns)
        return np.asarray(_eval_rows(alphas))
    pool = LapPool(track_json, width, vehicle_json, mode, ns, processes)
    try:
        return pool.lap_times(alphas)
    finally:
        pool.close()


# numpy dispatches `x ** 1.5` (path.py:58) to its vendored SVML `pow` on AVX512 hosts and to libm's elsewhere; the two
# differ by 1 ulp in ~5 % of the arguments and the reference's TBR18 lap times move by up to 6.5e-9 between such hosts.
NUMPY_AVX512_FEATURES = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR"
_SUBPROCESS_SNIPPET = """
import sys, numpy as np
sys.path.insert(0, {root!r})
from oracle.reference_port import lap_times_pool
np.save({out!r}, lap_times_pool({tj!r}, {width!r}, {vj!r}, np.load({inp!r}), {mode!r}, {ns!r}))
"""


def lap_times_baseline_dispatch(track_json, width, vehicle_json, alphas, mode="bayes", ns=None):
    """`lap_times_pool` in a process of its own with numpy's AVX512 kernels switched off (NPY_DISABLE_CPU_FEATURES is
    read when numpy is imported): the reference's arithmetic on a host without AVX512 (libm `pow`)."""
    import os
    import subprocess
    import sys
    import tempfile

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as d:
        inp, out = os.path.join(d, "a.npy"), os.path.join(d, "laps.npy")
        np.save(inp, np.asarray(alphas, dtype=np.float64))
        code = _SUBPROCESS_SNIPPET.format(root=root, inp=inp, out=out, tj=track_json, width=width, vj=vehicle_json,
                                          mode=mode, ns=ns)
        subprocess.run([sys.executable, "-c", code], env=dict(os.environ, NPY_DISABLE_CPU_FEATURES=NUMPY_AVX512_FEATURES),
                       check=True, capture_output=True)
        return np.load(out)

