"""`Path` facade -- same constructor, attributes and methods as the reference's `Path`
(src/path.py:17-77), evaluated by the CUDA spline kernel (`ltk_path_eval`, include/ltk.h).

The reference delegates to SciPy FITPACK (`splprep(k=3, s=0, per=1)`); the kernel builds the same
periodic interpolating cubic through the same knots as a cyclic tridiagonal system."""
from __future__ import annotations

import numpy as np

from . import _device


def cumulative_distances(points):
    """Cumulative chord length at each point, starting at 0 (path.py:11-14)."""
    return np.append(0, np.cumsum(np.linalg.norm(np.diff(points, axis=1), axis=0)))


class Path:
    """Periodic cubic spline through `controls` ([2, m], last column = first for closed paths),
    parameterised by cumulative chord length."""

    def __init__(self, controls, closed):
        self.controls = controls
        self.closed = closed
        self.dists = cumulative_distances(controls)
        self.length = self.dists[-1]
        if closed and isinstance(controls, np.ndarray):
            # splprep(per=1) closes the polygon IN PLACE on the caller's array (path.py:25); callers
            # depend on it (trajectory_bayesian_nonlinear.py:58-62), so the facade does the same.
            controls[:, -1] = controls[:, 0]
        self._dev = None

    # -- device evaluation -----------------------------------------------------------------------
    def _eval(self, u, want):
        if not self.closed:
            raise NotImplementedError("the CUDA path kernel handles closed paths only "
                                      "(open paths occur only in optimise_sectors; out of scope)")
        return _device.path_eval(self, np.atleast_1d(np.asarray(u, dtype=np.float64)), want)

    @property
    def spline(self):
        raise AttributeError("Path.spline (a FITPACK tck tuple in the reference) does not exist here: "
                             "the spline lives on the device; use position()/curvature()/gamma2()")

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
