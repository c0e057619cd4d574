#!/usr/bin/env python3
"""tests/golden/plateau_circle.npz: the UNMODIFIED reference on a circular corridor with regular cones and symmetric
candidates -- curvature plateaus, where np.argmin(v_local) (velocity.py:34,58) and the arg-max of the curvature name
different samples ("The sweeps' start sample").  Build container only (imports /root/reference/src).

    python tools/make_golden_plateau.py"""
import contextlib
import io
import json
import os
import sys
import tempfile
import warnings
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "plateau_circle.npz")
for m in ["casadi", "matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.colors"]:
    sys.modules[m] = MagicMock()
sys.path.insert(0, os.path.join(REF, "src"))
warnings.simplefilter("ignore")
with contextlib.redirect_stdout(io.StringIO()):
    from track import Track  # noqa: E402
    from vehicle import Vehicle  # noqa: E402
    from vehicleMX5 import VehicleMX5  # noqa: E402
    from trajectory import Trajectory  # noqa: E402
    from trajectory_bayesian_nonlinear import TrajectoryBayesianNonlinear  # noqa: E402

N_CONES, R_OUT, R_IN, WIDTH = 60, 50.0, 44.0, 0.8


def circle_doc():
    th = np.append(np.linspace(0.0, 2.0 * np.pi, N_CONES, endpoint=False), 0.0)
    return {"name": "circle", "left": {"x": list(R_OUT * np.cos(th)), "y": list(R_OUT * np.sin(th))},
            "right": {"x": list(R_IN * np.cos(th)), "y": list(R_IN * np.sin(th))}}


def population(n_alpha):
    rng = np.random.default_rng(1)
    return np.vstack([np.full((1, n_alpha), 0.5), np.full((1, n_alpha), 0.25), rng.uniform(0.45, 0.55, (30, n_alpha)),
                      np.tile(rng.uniform(0.0, 0.99, (32, 3)), (1, (n_alpha + 2) // 3))[:, :n_alpha]])


def main():
    out = {"n_cones": np.int64(N_CONES), "r_out": np.float64(R_OUT), "r_in": np.float64(R_IN), "width": np.float64(WIDTH)}
    with tempfile.TemporaryDirectory() as d:
        tj = os.path.join(d, "circle.json")
        with open(tj, "w") as fh:
            json.dump(circle_doc(), fh)
        for veh in ("tbr18", "mx5"):
            for mode in ("bayes", "full"):
                with contextlib.redirect_stdout(io.StringIO()):
                    track = Track(tj, track_width=WIDTH)
                    vehicle = VehicleMX5(f"{REF}/data/vehicles/MX5.json") if veh == "mx5" else Vehicle(f"{REF}/data/vehicles/tbr18.json")
                    T = TrajectoryBayesianNonlinear(track, vehicle) if mode == "bayes" else Trajectory(track, vehicle)
                n_alpha = len(track.mid_controls_decongested[0]) - int(track.closed) if mode == "bayes" else track.size
                a = population(n_alpha)
                laps, start, ties = np.empty(len(a)), np.empty(len(a), dtype=np.int64), np.empty(len(a), dtype=np.int64)
                for i, row in enumerate(a):
                    if mode == "bayes":
                        laps[i] = T.calcMinTime(T.updateAlphas(row.copy()))
                    else:
                        T.update(row.copy()); T.update_velocity(); laps[i] = T.lap_time()
                    vl = T.velocity.v_local
                    start[i] = int(np.argmin(vl))
                    k = T.path.curvature(T.s[:-1])
                    ties[i] = int(np.argmax(k)) != start[i]  # the two rules disagree on this candidate
                tag = f"{veh}_{mode}"
                out[tag + "_alphas"], out[tag + "_laps"], out[tag + "_start"], out[tag + "_rules_differ"] = a, laps, start, ties
                out[tag + "_ns"] = np.int64(T.ns)
                print(f"{tag}: {len(a)} candidates, n_alpha {n_alpha}, ns {T.ns}, laps {laps.min():.4f}..{laps.max():.4f}, "
                      f"arg-max(k) != arg-min(v_local) on {int(ties.sum())}")
    np.savez_compressed(OUT, **out)
    print("->", OUT, os.path.getsize(OUT), "B")


if __name__ == "__main__":
    main()
