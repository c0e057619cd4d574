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


def friction_circle(limit, lateral):
    """What is left of a friction circle of radius `limit` once `lateral` is spent sideways; nothing when the
    lateral demand alone reaches the limit (vehicle.py:33-35, vehicleMX5.py:35-37)."""
    return 0 if limit <= lateral else sqrt(limit**2 - lateral**2)


def announce(name, quiet):
    if not quiet:
        print("[ Imported {} ]".format(name))


class Vehicle:
    """Point mass with a tabulated engine map and a friction circle (TBR18)."""

    JSON_FIELDS = {"name": "name", "mass": "mass", "friction_coef": "frictionCoefficient"}

    def __init__(self, path, quiet=False):
        with open(path) as handle:
            doc = json.load(handle)
        for attr, key in self.JSON_FIELDS.items():
            setattr(self, attr, doc[key])
        table = doc["engineMap"]
        if len(table["v"]) > MAX_ENGINE_MAP:
            raise ValueError(f"engine map has more than {MAX_ENGINE_MAP} nodes")
        self.engine_profile = [table["v"], table["f"]]  # [speeds, forces] as in vehicle.py:18-21
        announce(self.name, quiet)

    def grip_limit(self):
        """(mu * m) * g, in the reference's order of multiplication (vehicle.py:30)."""
        return self.friction_coef * self.mass * GRAV

    def engine_force(self, velocity, gear=None):
        """Engine force at a speed: linear interpolation in the map, clamped (vehicle.py:25-27)."""
        speeds, forces = self.engine_profile
        return np.interp(velocity, speeds, forces)

    def traction(self, velocity, curvature):
        """Longitudinal force left inside the friction circle (vehicle.py:29-35)."""
        return friction_circle(self.grip_limit(), self.mass * velocity**2 * curvature)

    def to_ltk(self) -> LtkVehicle:
        """Constants for the kernels, folded in the reference's operation order."""
        out = LtkVehicle()
        out.kind, out.n_map = 0, len(self.engine_profile[0])
        out.mass = float(self.mass)
        out.mu_g = self.friction_coef * GRAV  # velocity.py:29
        out.f_max = self.grip_limit()
        out.f_max_sq = out.f_max**2
        for i, (speed, force) in enumerate(zip(*self.engine_profile)):
            out.map_v[i], out.map_f[i] = speed, force
        return out
