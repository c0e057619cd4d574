"""`Trajectory` facade (reference src/trajectory.py:26-58) plus the batched surface.

`update / update_velocity / lap_time` keep the reference's one-candidate call surface (each call runs
the CUDA pipeline for one candidate); `lap_time_batch` scores a whole population in one pipeline pass
-- the finite-difference gradient of `minimise_lap_time` (trajectory.py:128-146) is 132 such candidates."""
from __future__ import annotations

import math

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
