#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported read-only from
/root/reference/src with casadi/matplotlib stubbed -- SURVEY.md section 8(c)).

Runs only in the build container (the GPU box has no /root/reference).  The fixtures are committed;
tests compare (a) oracle/reference_port.py bit-for-bit, (b) oracle/lap_oracle.c and (c) the CUDA path
against them.

Two passes.  numpy evaluates `(dx**2 + dy**2) ** (3/2)` (path.py:58) with its vendored SVML pow on
AVX512 hosts (one ulp off the rounded value for ~5 % of the arguments) and with libm pow elsewhere, so the
reference's lap times depend on the host CPU at the 1e-10 level (TBR18).  Pass 1 records what the reference
produces here (AVX512 build container): `laps`, `prof_*`.  Pass 2 re-runs it in a child process with numpy's
AVX512 dispatch switched off (NPY_DISABLE_CPU_FEATURES) and adds `laps_base`, `prof_k_base`,
`prof_lap_base` -- the reference on numpy's baseline dispatch.  The spline internals of the profiled
candidates (`prof_tck_*`, `prof_d*`: splprep's tck and splev's derivatives) are the same in both.

Usage:  python tools/make_golden.py            # rewrites tests/golden/
"""
import io
import os
import sys
import contextlib
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")

for m in ["casadi", "matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.colors"]:
    sys.modules[m] = MagicMock()
sys.path.insert(0, os.path.join(REF, "src"))

import warnings  # noqa: E402

warnings.simplefilter("ignore")
with contextlib.redirect_stdout(io.StringIO()):
    from track import Track  # noqa: E402
    from vehicle import Vehicle  # noqa: E402
    from vehicleMX5 import VehicleMX5  # noqa: E402
    from trajectory import Trajectory  # noqa: E402
    from trajectory_bayesian_nonlinear import TrajectoryBayesianNonlinear  # noqa: E402

N_PROFILES = 4
BASE_PASS = os.environ.get("LTO_GOLDEN_BASE") == "1"
AVX512 = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR"


def make_vehicle(kind):
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "mx5":
            return VehicleMX5(f"{REF}/data/vehicles/MX5.json")
        return Vehicle(f"{REF}/data/vehicles/tbr18.json")


def make_track(name, width):
    with contextlib.redirect_stdout(io.StringIO()):
        return Track(f"{REF}/data/tracks/{name}.json", track_width=width)


def eval_bayes(T, a):
    lap = T.calcMinTime(T.updateAlphas(a))  # tbn.py:58-80
    return lap


def eval_full(T, a):
    T.update(a)  # trajectory.py:40-45
    T.update_velocity()  # :47-52
    return T.lap_time()  # :54-58


def grab(T, lap):
    from scipy.interpolate import splev

    vp = T.velocity
    k = T.path.curvature(T.s[:-1])
    t, c, _ = T.path.spline
    dx, dy = splev(T.s[:-1], T.path.spline, der=1)
    ddx, ddy = splev(T.s[:-1], T.path.spline, der=2)
    return dict(tck_t=np.array(t), tck_cx=np.array(c[0]), tck_cy=np.array(c[1]), dx=dx, dy=dy, ddx=ddx,
                ddy=ddy, dists=np.array(T.path.dists), lap=lap, s=T.s.copy(), k=np.array(k), v_local=vp.v_local.copy(),
                v_acclim=vp.v_acclim.copy(), v_declim=vp.v_declim.copy(), v=vp.v.copy(),
                controls=np.array(T.path.controls), length=T.path.length)


def run_case(tag, track_name, width, veh, mode, alphas, ns=None, n_profiles=N_PROFILES):
    track = make_track(track_name, width)
    vehicle = make_vehicle(veh)
    if mode == "bayes":
        T = TrajectoryBayesianNonlinear(track, vehicle)
        fn = eval_bayes
    else:
        T = Trajectory(track, vehicle)
        fn = eval_full
    if ns is not None:
        T.ns = ns
    laps = np.empty(len(alphas))
    prof = {}
    for i, a in enumerate(alphas):
        laps[i] = fn(T, a.copy())
        if i < n_profiles:
            for key, val in grab(T, laps[i]).items():
                prof.setdefault(key, []).append(val)
    out = dict(alphas=alphas, laps=laps, ns=np.int64(T.ns), track_length=np.float64(track.length),
               n_profiles=np.int64(min(n_profiles, len(alphas))))
    for key, val in prof.items():
        out["prof_" + key] = np.array(val)
    path = os.path.join(OUT, tag + ".npz")
    if BASE_PASS:
        first = dict(np.load(path))
        assert np.array_equal(first["alphas"], alphas)
        for key in ("tck_t", "tck_cx", "tck_cy", "dx", "dy", "ddx", "ddy", "s", "controls"):
            assert np.array_equal(first["prof_" + key], out["prof_" + key]), key
        first.update(laps_base=laps, prof_k_base=out["prof_k"], prof_lap_base=out["prof_lap"])
        out = first
    np.savez_compressed(path, **out)
    print(f"{tag}: {len(alphas)} candidates, ns={T.ns}, laps {laps.min():.4f}..{laps.max():.4f}"
          f" -> {os.path.getsize(path)} B")


def main():
    os.makedirs(OUT, exist_ok=True)
    n_bayes = {"buckmore": 43, "clay": 45, "gyg": 39, "whilton": 50}
    n_full = {"buckmore": 131, "clay": 137, "gyg": 119, "whilton": 151}
    rng = np.random.default_rng(20261018)

    def population(n, na, extra=True):
        a = rng.uniform(0.0, 0.99, (n, na))
        if extra and n >= 8:
            a[0] = 0.5  # centre line
            a[1] = 0.0
            a[2] = 0.99
            a[3] = rng.uniform(-0.46, 1.97, na)  # COBYLA leaves the box
            a[4] = np.where(np.arange(na) % 2 == 0, 0.05, 0.95)  # zig-zag: sharp curvature
        return a

    # the BASELINE.json configs, small
    run_case("buckmore_tbr18_bayes", "buckmore", 0.8, "tbr18", "bayes", population(256, 43))
    run_case("buckmore_mx5_bayes", "buckmore", 0.8, "mx5", "bayes", population(256, 43))
    run_case("buckmore_tbr18_full", "buckmore", 0.8, "tbr18", "full", population(96, 131))
    run_case("buckmore_mx5_full", "buckmore", 0.8, "mx5", "full", population(96, 131))
    # other tracks / widths
    for t in ["clay", "gyg", "whilton"]:
        run_case(f"{t}_tbr18_bayes", t, 0.8, "tbr18", "bayes", population(48, n_bayes[t]))
        run_case(f"{t}_mx5_full", t, 0.6, "mx5", "full", population(16, n_full[t]))
    run_case("buckmore_w100_tbr18_bayes", "buckmore", 1.0, "tbr18", "bayes", population(32, 43))
    # ns override (BASELINE.json config 5: densely resampled lap)
    run_case("buckmore_tbr18_bayes_ns2501", "buckmore", 0.8, "tbr18", "bayes",
             population(12, 43), ns=2501, n_profiles=2)
    run_case("buckmore_tbr18_bayes_ns10001", "buckmore", 0.8, "tbr18", "bayes",
             population(6, 43, extra=False), ns=10001, n_profiles=1)
    # SURVEY.md known-answer candidate: default_rng(0).uniform(0, 0.99, 43) -> 45.16138534803076
    ka = np.random.default_rng(0).uniform(0, 0.99, (1, 43))
    run_case("buckmore_tbr18_bayes_known", "buckmore", 0.8, "tbr18", "bayes", ka)


if __name__ == "__main__":
    main()
    if not BASE_PASS:
        import subprocess

        env = dict(os.environ, LTO_GOLDEN_BASE="1", NPY_DISABLE_CPU_FEATURES=AVX512)
        subprocess.check_call([sys.executable, os.path.abspath(__file__)], env=env)
