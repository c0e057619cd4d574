"""`Trajectory` facade (reference src/trajectory.py:26-58) plus the batched surface.

`update / update_velocity / lap_time` keep the reference's one-candidate call surface (each call runs
the CUDA pipeline for one candidate); `lap_time_batch` scores a whole population in one pipeline pass
-- the finite-difference gradient of `minimise_lap_time` (trajectory.py:128-146) is 132 such candidates."""
from __future__ import annotations

import math
import time

import numpy as np

from .evaluator import LapTimeEvaluator
from .path import Path
from .velocity import VelocityProfile


class Trajectory:
    """Geometry and dynamics of one racing line; samples are taken every metre of centre line."""

    MODE = "full"

    def __init__(self, track, vehicle, device=None):
        self.track = track
        self.ns = math.ceil(track.length)
        self.vehicle = vehicle
        self.velocity = None
        self._device = device
        self._evaluator = None
        self._lap = None
        self.update(np.full(track.size, 0.5))

    # -- device evaluator, re-created when ns changes ---------------------------------------------
    @property
    def evaluator(self) -> LapTimeEvaluator:
        if self._evaluator is None:
            self._evaluator = LapTimeEvaluator(self.track, self.vehicle, self.MODE, self.ns, self._device)
        elif self._evaluator.ns != self.ns:
            self._evaluator.set_ns(self.ns)
        return self._evaluator

    # -- reference surface ------------------------------------------------------------------------
    def update(self, alphas):
        """New control points and path (trajectory.py:40-45)."""
        self.alphas = alphas
        self.path = Path(self.track.control_points(alphas), self.track.closed)
        self.s = np.linspace(0, self.path.length, self.ns)
        self._lap = None

    def update_velocity(self):
        """Velocity profile of the current path (trajectory.py:47-52), one pipeline pass on the GPU."""
        prof = self.evaluator.profile(np.asarray(self.alphas, dtype=np.float64))
        s_max = self.path.length if self.track.closed else None
        self.velocity = VelocityProfile(self.vehicle, self.s[:-1], prof["k"], s_max,
                                        _precomputed=(prof["v_local"], prof["v_acclim"], prof["v_declim"], prof["v"]))
        self._lap = prof["lap"]

    def lap_time(self):
        """Lap time of the current velocity profile (trajectory.py:54-58); summed on the device."""
        if self._lap is None:
            self.update_velocity()
        return self._lap

    # -- batched surface --------------------------------------------------------------------------
    def lap_time_batch(self, alphas):
        """alphas [B, track.size] (numpy or CUDA tensor) -> lap times [B] of the same kind."""
        if isinstance(alphas, np.ndarray) or not hasattr(alphas, "is_cuda"):
            return self.evaluator.lap_times(alphas)
        return self.evaluator.lap_times_device(alphas)

    def curvature_objectives_batch(self, alphas):
        """alphas [B, track.size] -> (gamma2[B], length[B]): `path.gamma2(self.s)` and `path.length`
        (trajectory.py:60-97) of every candidate, one pipeline pass."""
        return self.evaluator.curvature_objectives(alphas)

    # -- optimisers: the reference's objectives with batched finite-difference gradients -------------
    # The reference hands `objfun` to scipy's L-BFGS-B without a gradient (trajectory.py:60-146), so scipy
    # differences it itself: N + 1 = 132 sequential evaluations per gradient.  Here the same 2-point
    # scheme (absolute step 1e-8, flipped where it would leave the box -- scipy's `approx_derivative`
    # with `abs_step` and bounds) is ONE batch of N + 1 candidates on the GPU.
    FD_STEP = 1e-8

    def _fd_points(self, x, lo=0.0, hi=1.0):
        x = np.asarray(x, dtype=np.float64)
        h = np.full(x.size, self.FD_STEP)
        h[x + h > hi] *= -1.0
        pts = np.repeat(x[None, :], x.size + 1, axis=0)
        idx = np.arange(x.size)
        pts[idx + 1, idx] = x + h
        dx = pts[idx + 1, idx] - x  # the step actually taken, as scipy does
        return pts, dx

    def _minimise(self, batch_objective):
        from scipy.optimize import Bounds, minimize

        def fun(x):
            pts, dx = self._fd_points(x)
            f = batch_objective(pts)
            return float(f[0]), (f[1:] - f[0]) / dx

        t0 = time.time()
        res = minimize(fun=fun, x0=np.full(self.track.size, 0.5), jac=True, method="L-BFGS-B", bounds=Bounds(0.0, 1.0))
        self.update(res.x)
        self.result = res
        return time.time() - t0

    def lap_time_and_gradient(self, alphas):
        """Lap time and its forward-difference gradient at `alphas`: one batch of N + 1 candidates."""
        pts, dx = self._fd_points(alphas)
        f = self.evaluator.lap_times(pts)
        return f[0], (f[1:] - f[0]) / dx

    def minimise_curvature(self):
        """Generate a path minimising curvature (trajectory.py:60-75)."""
        return self._minimise(lambda pts: self.evaluator.curvature_objectives(pts)[0])

    def minimise_compromise(self, eps):
        """Compromise between curvature and path length, `eps` weighs the length (trajectory.py:77-97)."""

        def obj(pts):
            k, d = self.evaluator.curvature_objectives(pts)
            return (1 - eps) * k + eps * d

        return self._minimise(obj)

    def minimise_optimal_compromise(self, eps_min=0, eps_max=0.2):
        """The compromise weight with the lowest lap time: a bounded scalar search over `eps`, each probe
        one `minimise_compromise` run followed by one lap-time evaluation of its path; leaves `epsilon`,
        `epsilon_history` ([eps, lap time] rows, one row stays one-dimensional as in the reference) and
        the path of the best weight (trajectory.py:99-126)."""
        from scipy.optimize import minimize_scalar

        probes = []

        def fun(eps):
            self.minimise_compromise(eps)
            lap = float(self.evaluator.lap_times(self.alphas)[0])
            probes.append([eps, lap])
            return lap

        t0 = time.time()
        res = minimize_scalar(fun=fun, method="bounded", bounds=(eps_min, eps_max))
        self.epsilon_history = np.array(probes[0]) if len(probes) == 1 else np.array(probes)
        self.epsilon = res.x
        self.minimise_compromise(self.epsilon)
        return time.time() - t0

    def minimise_lap_time(self):
        """Generate a path that directly minimises lap time (trajectory.py:128-146)."""
        return self._minimise(self.evaluator.lap_times)
