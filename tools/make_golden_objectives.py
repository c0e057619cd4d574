#!/usr/bin/env python3
"""Golden vectors for the curvature / path-length objectives of `Trajectory.minimise_curvature` and
`minimise_compromise` (reference src/trajectory.py:60-97: `path.gamma2(self.s)`, `path.length`), produced
by the UNMODIFIED reference imported from /root/reference/src (casadi/matplotlib stubbed).

    python tools/make_golden_objectives.py      # writes tests/golden/objectives_*.npz

Runs only in the build container; the fixtures are committed."""
import contextlib
import io
import os
import sys
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
    for name, width, seed, count in (("buckmore", 0.8, 4101, 48), ("clay", 0.8, 4102, 24)):
        with contextlib.redirect_stdout(io.StringIO()):
            track = Track(f"{REF}/data/tracks/{name}.json", track_width=width)
            T = Trajectory(track, Vehicle(f"{REF}/data/vehicles/tbr18.json"))
        rng = np.random.default_rng(seed)
        alphas = rng.uniform(0.0, 1.0, (count, track.size))
        alphas[0] = 0.5  # the centre line, the optimisers' starting point (trajectory.py:69)
        alphas[1] = np.clip(0.5 + 0.4 * np.sin(np.arange(track.size) * 0.3), 0, 1)
        g2, length = [], []
        for a in alphas:
            T.update(a)                       # trajectory.py:40-45
            g2.append(T.path.gamma2(T.s))     # path.py:63-77 at all ns samples, end point included
            length.append(T.path.length)
        np.savez_compressed(os.path.join(OUT, f"objectives_{name}_full.npz"), alphas=alphas, gamma2=np.array(g2),
                            length=np.array(length), ns=np.int64(T.ns), width=np.float64(width))
        print(name, "ns", T.ns, "gamma2", g2[0], "length", length[0])


if __name__ == "__main__":
    main()
