"""lap_time_optimization_b200 -- B200-native batched lap-time evaluation behind the reference's
Path / VelocityProfile / Trajectory call surface (bruno-maruszczak/lap-time-optimization).

Importing the package needs neither a GPU nor the built library; anything that computes loads
`libltk.so` and raises `LtkUnavailable` if it (or a CUDA device) is missing -- there is no CPU fallback."""
import os

from ._native import LtkError, LtkUnavailable  # noqa: F401
from .track import Track  # noqa: F401
from .vehicle import Vehicle  # noqa: F401
from .vehicleMX5 import VehicleMX5  # noqa: F401
from .path import Path, default_spline, set_default_spline  # noqa: F401
from .velocity import VelocityProfile  # noqa: F401
from .evaluator import LapTimeEvaluator  # noqa: F401
from .trajectory import Trajectory  # noqa: F401
from .trajectory_bayesian_nonlinear import TrajectoryBayesianNonlinear  # noqa: F401

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def data_path(*parts):
    """Path of a bundled input file, e.g. data_path("tracks", "buckmore.json")."""
    return os.path.join(DATA_DIR, *parts)


def load_vehicle(path, quiet=True):
    """The reference picks the class by comparing the path string (__main__.py:100); we look inside."""
    with open(path) as f:
        text = f.read()
    return VehicleMX5(path, quiet=quiet) if '"control"' in text else Vehicle(path, quiet=quiet)
