"""`VelocityProfile` facade -- the reference's three-pass profile (src/velocity.py:9-76) computed by
the CUDA kernel `ltk_velocity_profile` for caller-supplied samples.  The batched pipeline
(`LapTimeEvaluator`) fuses the same arithmetic into the K23 sweep kernel."""
from __future__ import annotations

import numpy as np

from . import _device

GRAV = 9.81  # ms^-2 (velocity.py:4)


class VelocityProfile:
    """`s` and `k` exclude the overlapping end sample of a closed path; `s_max` is the lap length of
    a closed path or None for an open one (velocity.py:14-26)."""

    def __init__(self, vehicle, s, k, s_max=None, _precomputed=None):
        self.vehicle = vehicle
        self.s = s
        self.s_max = s_max
        if _precomputed is not None:  # filled by Trajectory from one ltk_profile call
            self.v_local, self.v_acclim, self.v_declim, self.v = _precomputed
            return
        self._run(k)

    def _run(self, k):
        self.v_local, self.v_acclim, self.v_declim, self.v = _device.velocity_profile(
            self.vehicle, np.asarray(self.s, dtype=np.float64), np.asarray(k, dtype=np.float64), self.s_max)

    # the reference exposes its three passes as methods; each refreshes the attribute it owns
    def limit_local_velocities(self, k):
        self._run(k)

    def limit_acceleration(self, k_in):
        self._run(k_in)

    def limit_deceleration(self, k_in):
        self._run(k_in)
