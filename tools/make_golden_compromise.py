#!/usr/bin/env python3
"""Golden result of `Trajectory.minimise_optimal_compromise` (reference src/trajectory.py:99-126) from the
UNMODIFIED reference imported from /root/reference/src (casadi/matplotlib stubbed): the probes of the
bounded scalar search over the compromise weight (`epsilon_history`), the chosen weight and the lap time of
its path, and the curvature / length objectives every inner `minimise_compromise` run ended on.

    python tools/make_golden_compromise.py      # writes tests/golden/compromise_buckmore.npz  (minutes)

Runs only in the build container; the fixture is committed."""
import contextlib
import io
import os
import sys
import time
import warnings
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
for m in ["casadi", "matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.colors"]:
    sys.modules[m] = MagicMock()
sys.path.insert(0, os.path.join(REF, "src"))
warnings.simplefilter("ignore")
with contextlib.redirect_stdout(io.StringIO()):
    from track import Track  # noqa: E402
    from vehicle import Vehicle  # noqa: E402
    from trajectory import Trajectory  # noqa: E402


def main():
    width = 0.8
    with contextlib.redirect_stdout(io.StringIO()):
        track = Track(f"{REF}/data/tracks/buckmore.json", track_width=width)
        T = Trajectory(track, Vehicle(f"{REF}/data/vehicles/tbr18.json"))
    reached = []                              # what every inner L-BFGS-B run ended on
    inner = T.minimise_compromise

    def logged(eps):
        spent = inner(eps)                    # trajectory.py:77-97, unmodified
        reached.append([eps, T.path.gamma2(T.s), T.path.length])
        return spent

    T.minimise_compromise = logged
    t0 = time.time()
    T.minimise_optimal_compromise()           # trajectory.py:99-126, defaults eps in [0, 0.2]
    T.update_velocity()
    lap = T.lap_time()
    print(f"{time.time() - t0:.1f} s; epsilon {T.epsilon:.6f}; lap {lap:.6f}; probes:\n{T.epsilon_history}")
    np.savez_compressed(os.path.join(OUT, "compromise_buckmore.npz"), epsilon=np.float64(T.epsilon),
                        history=np.atleast_2d(T.epsilon_history), lap=np.float64(lap), alphas=np.asarray(T.alphas),
                        reached=np.array(reached),  # [eps, sum of squared curvatures, path length] per inner run
                        width=np.float64(width), ns=np.int64(T.ns))


if __name__ == "__main__":
    main()
