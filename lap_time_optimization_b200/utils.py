"""Small host helpers the hot-path callers use (reference src/utils.py:5-22)."""
import numpy as np


def idx_modulo(a, b, n):
    """Indices a..b-1 on a ring of n elements (utils.py:5-14)."""
    i, j = a % n, b % n
    if i < j:
        return np.arange(i, j, dtype=int)
    return np.append(np.arange(i, n, dtype=int), np.arange(0, j, dtype=int))


def is_closed(left, right):
    """A track is a closed loop when both boundaries end where they start (utils.py:17-22)."""
    return bool(all(left[:, 0] == left[:, -1]) and all(right[:, 0] == right[:, -1]))


# ---- result artefacts: the JSON files the MPC demo consumes (utils.py:108-136, __main__.py:196-213) ----
def _dump(directory, name, data):
    import json
    import os

    os.makedirs(directory, exist_ok=True)
    # the reference joins with a literal backslash (f"{path}\{name}.json", utils.py:117), which only
    # works on Windows; the format of the file is what mpc/track.py:15-19,44-57 reads
    with open(os.path.join(directory, name + ".json"), "w") as f:
        json.dump(data, f, indent=4)


def save_path_to_json(path, x, y, name):
    """`{"name", "path": {"x": [...], "y": [...]}}` (utils.py:108-118)."""
    _dump(path, name, {"name": name, "path": {"x": np.asarray(x).tolist(), "y": np.asarray(y).tolist()}})


def save_widths_to_json(path, x, name):
    """`{"name", "width": [...]}` (utils.py:120-127)."""
    _dump(path, name, {"name": name, "width": np.asarray(x).tolist()})


def save_velocities_to_json(path, x, name):
    """`{"name", "velocities": [...]}` (utils.py:129-136)."""
    _dump(path, name, {"name": name, "velocities": np.asarray(x).tolist()})


def save_result_artefacts(directory, track, trajectory):
    """The five files `src/__main__.py:196-213` leaves next to its plots for the MPC demo: the sampled racing
    line, both (unshrunk) boundaries, the track widths and the velocity profile of `trajectory`."""
    xy = trajectory.path.position(trajectory.s)
    save_path_to_json(directory, xy[0], xy[1], "path")
    save_path_to_json(directory, track.old_left[0], track.old_left[1], "left")
    save_path_to_json(directory, track.old_right[0], track.old_right[1], "right")
    save_widths_to_json(directory, track.widths, "widths")
    save_velocities_to_json(directory, trajectory.velocity.v, "velocities")
