"""MX-5 vehicle -- host side.  Mirrors the reference's `VehicleMX5` (src/vehicleMX5.py:10-79): a
Pacejka-parameterised car reduced to the two callbacks the velocity profile needs."""
from __future__ import annotations

import json
import re
from math import sqrt

from ._native import LtkVehicle

GRAV = 9.81  # m/s^2 (vehicleMX5.py:6)


def _strip_json_comments(text):
    text = re.sub(r"//.*", "", text)
    return re.sub(r"/\*.*?\*/", "", text, flags=re.DOTALL)


class VehicleMX5:
    def __init__(self, vehicle_filepath, quiet=False):
        self.load_params(vehicle_filepath)
        if not quiet:
            print("[ Imported {} ]".format(self.name))

    def engine_force(self, velocity, gear=None):
        """Maximum longitudinal drive force (vehicleMX5.py:19-21)."""
        return (self.T * self.C_m) - self.Cr_0 - (self.Cr_2 * (velocity**2))

    def traction(self, v, k, lam=2.0):
        """Force left inside the friction circle of radius lam*D*m*g (vehicleMX5.py:23-37)."""
        D = (self.D_f + self.D_r) * 0.5
        Fn = self.mass * GRAV
        F_max = lam * D * Fn
        F_lat = self.mass * v * v * k
        if F_max <= F_lat:
            return 0
        return sqrt(F_max**2 - F_lat**2)

    def remove_comments(self, json_str):
        return _strip_json_comments(json_str)

    def load_params(self, vehicle_filepath):
        """Field names follow vehicleMX5.py:46-79."""
        with open(vehicle_filepath) as f:
            data = json.loads(_strip_json_comments(f.read()))
        self.rotational_inertia = data["rotational_inertia"]
        self.name = data["name"]
        self.mass = data["mass"]
        self.length_f = data["length_f"]
        self.length_r = data["length_r"]
        self.B_f, self.C_f, self.D_f = (data["frontTire"][k] for k in ("B_f", "C_f", "D_f"))
        self.B_r, self.C_r, self.D_r = (data["rearTire"][k] for k in ("B_r", "C_r", "D_r"))
        self.C_m = data["control"]["C_m"]
        self.Cr_0 = data["Cr_0"]
        self.Cr_2 = data["Cr_2"]
        self.ptv = data["ptv"]
        self.T = data["control"]["T"]
        self.friction_coef = data["control"]["lambda"]
        self.ro_long = data["control"]["ro_long"]

    def to_ltk(self, lam=2.0) -> LtkVehicle:
        v = LtkVehicle()
        v.kind = 1
        v.n_map = 0
        v.mass = float(self.mass)
        v.mu_g = self.friction_coef * GRAV  # velocity.py:29 with friction_coef = control.lambda
        D = (self.D_f + self.D_r) * 0.5
        Fn = self.mass * GRAV
        f = lam * D * Fn
        v.f_max = f
        v.f_max_sq = f**2
        v.e0 = (self.T * self.C_m) - self.Cr_0
        v.cr2 = self.Cr_2
        return v
