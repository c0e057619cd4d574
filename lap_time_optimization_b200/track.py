"""`Track` -- host side, runs once per problem.  Mirrors the reference's `Track` (src/track.py:8-118):
reads cone pairs, shrinks the corridor to a width fraction, exposes the per-control-point affine map
alpha -> control point that the K1 kernel applies on the device."""
from __future__ import annotations

import json

import numpy as np

from .path import Path
from .utils import is_closed

DECONGEST_STRIDE = 3  # track.py:40


class Track:
    """Track boundaries as two cone polylines, `left[:, i]` paired with `right[:, i]`."""

    def __init__(self, json_path=None, left=None, right=None, track_width=None, quiet=False):
        self._quiet = quiet
        if json_path is not None:
            # the stored value is the fraction REMOVED from the corridor (track.py:17-21)
            self.track_width = 1.0 - min(max(track_width, 0.001), 1.0)
            self.read_cones(json_path)
        else:
            self.left, self.right = left, right
        self._derive_geometry()
        self._decongest()

    def _derive_geometry(self):
        """closed / size / diffs / centre line / length (track.py:23-28)."""
        self.closed = is_closed(self.left, self.right)
        self.size = self.left[0].size - int(self.closed)
        self.diffs = self.right - self.left
        self.mid = Path(self.control_points(np.full(self.size, 0.5)), self.closed)
        self.length = self.mid.dists[-1]
        self.widths = np.sqrt(self.diffs[0] ** 2 + self.diffs[1] ** 2)
        self.mid_controls = self.mid.controls

    def _decongest(self):
        """The every-3rd-cone subset the Bayesian / nonlinear optimisers work on (track.py:31-49)."""
        pick = np.arange(0, self.mid.controls.shape[1], DECONGEST_STRIDE)
        self.mid_controls_decongested = [list(row[pick]) for row in self.mid.controls]
        self.widths_decongested = list(self.widths[pick])
        self.left_decongested, self.diffs_decongested = self.left[:, pick], self.diffs[:, pick]

    def read_cones(self, path):
        """Cone coordinates from `{"name", "left": {"x","y"}, "right": {"x","y"}}` (track.py:52-70); the
        boundaries used from here on are the ones pulled in to the requested width."""
        with open(path) as handle:
            doc = json.load(handle)
        self.name = doc["name"]
        if not self._quiet:
            print("[ Imported {} ]".format(self.name))
        self.old_left, self.old_right = (np.array([doc[side]["x"], doc[side]["y"]]) for side in ("left", "right"))
        self.new_left = self.new_left_cones(self.old_left, self.old_right, self.track_width)
        self.new_right = self.new_right_cones(self.old_left, self.old_right, self.track_width)
        self.left, self.right = self.new_left, self.new_right

    def avg_curvature(self, s):
        """Mean centre-line curvature at the sample distances (track.py:73-76)."""
        return np.sum(self.mid.curvature(s)) / s.size

    def _place(self, origin, span, alphas):
        """origin + alpha * span per control point; a closed track repeats the first alpha at the end, and
        entries equal to -1 are skipped (track.py:82-94)."""
        alphas = np.asarray(alphas, dtype=float)
        if self.closed:
            alphas = np.append(alphas, alphas[0])
        keep = np.nonzero(alphas != -1)[0]
        return origin[:, keep] + (alphas[keep] * span[:, keep])

    def control_points(self, alphas):
        """alpha in [0,1] per cone -> point on the segment left->right (track.py:82-87)."""
        return self._place(self.left, self.diffs, alphas)

    def control_points_bayesian(self, alphas):
        """Same map on the every-3rd-cone subset (track.py:89-94)."""
        return self._place(self.left_decongested, self.diffs_decongested, alphas)

    @staticmethod
    def _shrink(inner, outer, fraction):
        """Move `inner` towards `outer` by fraction/2 of the cone-to-cone vector (track.py:96-118)."""
        inner, outer = np.asarray(inner, dtype=float), np.asarray(outer, dtype=float)
        return inner + fraction * (outer - inner) / 2

    def new_left_cones(self, old_left, old_right, track_width):
        return self._shrink(old_left, old_right, track_width)

    def new_right_cones(self, old_left, old_right, track_width):
        return self._shrink(old_right, old_left, track_width)

    # -- constants for the device kernels ---------------------------------------------------------
    def affine_map(self, mode):
        """(left_xy, diff_xy), each [2, N] over the N UNIQUE control points of a closed track.
        mode "full": one alpha per cone (Trajectory, trajectory.py:40-45);
        mode "bayes": every-3rd-cone subset; the last subset cone is dropped because the reference's
        closure overwrites it with the first."""
        if not self.closed:
            raise NotImplementedError("batched evaluation supports closed tracks only")
        if mode == "full":
            l, d = self.left, self.diffs
        elif mode == "bayes":
            l, d = self.left_decongested, self.diffs_decongested
        else:
            raise ValueError("mode must be 'full' or 'bayes'")
        return np.ascontiguousarray(l[:, :-1]), np.ascontiguousarray(d[:, :-1])
