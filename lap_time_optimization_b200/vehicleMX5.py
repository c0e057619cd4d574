"""MX-5 vehicle -- host side.  Mirrors the reference's `VehicleMX5` (src/vehicleMX5.py:10-79): a
Pacejka-parameterised car reduced to the two callbacks the velocity profile needs."""
from __future__ import annotations

import json
import re

from ._native import LtkVehicle
from .vehicle import GRAV, announce, friction_circle

# attribute -> where it sits in the (commented) JSON document; the names are the ones callers of the
# reference read (vehicleMX5.py:46-79)
JSON_FIELDS = {
    "rotational_inertia": ("rotational_inertia",), "name": ("name",), "mass": ("mass",),
    "length_f": ("length_f",), "length_r": ("length_r",),
    "B_f": ("frontTire", "B_f"), "C_f": ("frontTire", "C_f"), "D_f": ("frontTire", "D_f"),
    "B_r": ("rearTire", "B_r"), "C_r": ("rearTire", "C_r"), "D_r": ("rearTire", "D_r"),
    "C_m": ("control", "C_m"), "Cr_0": ("Cr_0",), "Cr_2": ("Cr_2",), "ptv": ("ptv",),
    "T": ("control", "T"), "friction_coef": ("control", "lambda"), "ro_long": ("control", "ro_long"),
}


def _strip_json_comments(text):
    text = re.sub(r"//.*", "", text)
    return re.sub(r"/\*.*?\*/", "", text, flags=re.DOTALL)


class VehicleMX5:
    def __init__(self, vehicle_filepath, quiet=False):
        self.load_params(vehicle_filepath)
        announce(self.name, quiet)

    def remove_comments(self, json_str):
        return _strip_json_comments(json_str)

    def load_params(self, vehicle_filepath):
        with open(vehicle_filepath) as handle:
            doc = json.loads(_strip_json_comments(handle.read()))
        for attr, keys in JSON_FIELDS.items():
            node = doc
            for key in keys:
                node = node[key]
            setattr(self, attr, node)

    def drive_offset(self):
        """(T * C_m) - Cr_0: the speed-independent part of the drive force (vehicleMX5.py:21)."""
        return (self.T * self.C_m) - self.Cr_0

    def grip_limit(self, lam=2.0):
        """lam * D * (m * g) with D the mean Pacejka peak factor of the two axles (vehicleMX5.py:28-33)."""
        peak = (self.D_f + self.D_r) * 0.5
        return lam * peak * (self.mass * GRAV)

    def engine_force(self, velocity, gear=None):
        """Maximum longitudinal drive force (vehicleMX5.py:19-21)."""
        return self.drive_offset() - (self.Cr_2 * (velocity**2))

    def traction(self, v, k, lam=2.0):
        """Force left inside the friction circle of radius lam*D*m*g (vehicleMX5.py:23-37)."""
        return friction_circle(self.grip_limit(lam), self.mass * v * v * k)

    def to_ltk(self, lam=2.0) -> LtkVehicle:
        out = LtkVehicle()
        out.kind, out.n_map = 1, 0
        out.mass = float(self.mass)
        out.mu_g = self.friction_coef * GRAV  # velocity.py:29 with friction_coef = control.lambda
        out.f_max = self.grip_limit(lam)
        out.f_max_sq = out.f_max**2
        out.e0 = self.drive_offset()
        out.cr2 = self.Cr_2
        return out
