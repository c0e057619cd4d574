"""`Path` facade -- same constructor, attributes and methods as the reference's `Path`
(src/path.py:17-77), evaluated by the CUDA spline kernels (`ltk_path_eval` / `ltk_path_eval_fitpack`,
include/ltk.h).

The reference delegates to SciPy FITPACK (`splprep(k=3, s=0, per=closed)`).  Two arithmetics build that spline
here (`lap_time_optimization_b200.set_default_spline`, or the `spline=` argument):
  "tridiagonal" -- the same periodic interpolating cubic through the same knots as a cyclic tridiagonal system
                   (closed paths only; curvature within ~1e-13 of SciPy's);
  "fitpack"     -- FITPACK's own algorithm and operation order (fpclos / fppara Givens QR, splder, fpbspl): knots,
                   coefficients, positions and derivatives bit-equal to SciPy's, closed and open paths.
`Path.spline` (the tck tuple the reference keeps) always comes from the second one."""
from __future__ import annotations

import numpy as np

from . import _device


def cumulative_distances(points):
    """Cumulative chord length at each point, starting at 0 (path.py:11-14)."""
    return np.append(0, np.cumsum(np.linalg.norm(np.diff(points, axis=1), axis=0)))


_DEFAULT = {"spline": "tridiagonal"}


def set_default_spline(mode):
    """Spline arithmetic used by `Path` objects and `LapTimeEvaluator`s created from now on that do not name one:
    "tridiagonal" (fastest) or "fitpack" (SciPy FITPACK's own bits; see include/ltk.h LTK_SPLINE_*)."""
    if mode not in ("tridiagonal", "fitpack"):
        raise ValueError("spline mode must be 'tridiagonal' or 'fitpack'")
    _DEFAULT["spline"] = mode


def default_spline():
    return _DEFAULT["spline"]


class Path:
    """Interpolating cubic spline through `controls` ([2, m]; closed paths: periodic, last column = first),
    parameterised by cumulative chord length."""

    def __init__(self, controls, closed, spline=None):
        self.controls = controls
        self.closed = closed
        self._spline_mode = spline
        self._tck = None
        self.dists = cumulative_distances(controls)
        self.length = self.dists[-1]
        if closed and isinstance(controls, np.ndarray):
            # splprep(per=1) closes the polygon IN PLACE on the caller's array (path.py:25); callers
            # depend on it (trajectory_bayesian_nonlinear.py:58-62), so the facade does the same.
            controls[:, -1] = controls[:, 0]
        self._dev = None

    # -- device evaluation -----------------------------------------------------------------------
    def _eval(self, u, want):
        u = np.atleast_1d(np.asarray(u, dtype=np.float64))
        mode = self._spline_mode or _DEFAULT["spline"]
        if mode == "fitpack" or not self.closed:  # the tridiagonal kernel is periodic-only
            return _device.path_eval_fitpack(self, u, want)
        return _device.path_eval(self, u, want)

    @property
    def spline(self):
        """The tck tuple `(t, [cx, cy], 3)` the reference keeps in `Path.spline` (path.py:25): knots and B-spline
        coefficients bit-equal to `splprep(controls, u=dists, k=3, s=0, per=closed)[0]`, computed on the GPU."""
        if self._tck is None:
            out = _device.path_eval_fitpack(self, np.empty(0), ("tck",))
            self._tck = (out["t"], [out["cx"], out["cy"]], 3)
        return self._tck

    def position(self, s=None):
        """x-y coordinates at parameters s (path.py:29-34)."""
        if s is None:
            return self.controls
        out = self._eval(s, ("x", "y"))
        return np.array([out["x"], out["y"]])

    def curvature(self, u=None, return_absolute_value=True):
        """Curvature (x'y'' - y'x'') / (x'^2 + y'^2)^(3/2) at parameters u (path.py:36-61)."""
        if u is None:
            u = self.dists
        k = self._eval(u, ("k",))["k"]
        return np.abs(k) if return_absolute_value else k

    def gamma2(self, u=None):
        """Sum of squared curvatures at the samples (path.py:63-77)."""
        if u is None:
            u = self.dists
        return self._eval(u, ("gamma2",))["gamma2"]
