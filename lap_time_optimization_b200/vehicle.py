"""Vehicle description -- host side.  Mirrors the reference's `Vehicle` (src/vehicle.py:11-35): same
constructor, attributes (`name`, `mass`, `friction_coef`, `engine_profile`) and scalar callbacks.  The
callbacks are plain host arithmetic for callers that poke at one value; the lap-time path itself runs
in the CUDA kernels, which receive the constants through `to_ltk()`."""
from __future__ import annotations

import json
from math import sqrt

import numpy as np

from ._native import MAX_ENGINE_MAP, LtkVehicle

GRAV = 9.81  # m/s^2 (vehicle.py:5)


class Vehicle:
    """Point mass with a tabulated engine map and a friction circle (TBR18)."""

    def __init__(self, path, quiet=False):
        with open(path) as f:
            data = json.load(f)
        self.name = data["name"]
        self.mass = data["mass"]
        self.friction_coef = data["frictionCoefficient"]
        self.engine_profile = [data["engineMap"]["v"], data["engineMap"]["f"]]
        if len(self.engine_profile[0]) > MAX_ENGINE_MAP:
            raise ValueError(f"engine map has more than {MAX_ENGINE_MAP} nodes")
        if not quiet:
            print("[ Imported {} ]".format(self.name))

    def engine_force(self, velocity, gear=None):
        """Engine force at a speed: linear interpolation in the map, clamped (vehicle.py:25-27)."""
        return np.interp(velocity, self.engine_profile[0], self.engine_profile[1])

    def traction(self, velocity, curvature):
        """Longitudinal force left inside the friction circle (vehicle.py:29-35)."""
        f = self.friction_coef * self.mass * GRAV
        f_lat = self.mass * velocity**2 * curvature
        if f <= f_lat:
            return 0
        return sqrt(f**2 - f_lat**2)

    def to_ltk(self) -> LtkVehicle:
        """Constants for the kernels, folded in the reference's operation order."""
        v = LtkVehicle()
        v.kind = 0
        v.n_map = len(self.engine_profile[0])
        v.mass = float(self.mass)
        v.mu_g = self.friction_coef * GRAV  # velocity.py:29
        f = self.friction_coef * self.mass * GRAV  # vehicle.py:30
        v.f_max = f
        v.f_max_sq = f**2
        for i, (x, y) in enumerate(zip(*self.engine_profile)):
            v.map_v[i] = x
            v.map_f[i] = y
        return v
