import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def golden_cases():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    # objectives_*: tools/make_golden_objectives.py, compromise_*: tools/make_golden_compromise.py
    # plateau_*: tools/make_golden_plateau.py (a synthetic circular corridor, generated on the fly by its tests)
    return [n for n in names if not n.startswith(("objectives_", "compromise_", "plateau_"))]


def circle_track_json(directory, g):
    """The circular corridor of tests/golden/plateau_circle.npz (tools/make_golden_plateau.py::circle_doc)."""
    import json

    th = np.append(np.linspace(0.0, 2.0 * np.pi, int(g["n_cones"]), endpoint=False), 0.0)
    doc = {"name": "circle", "left": {"x": list(float(g["r_out"]) * np.cos(th)), "y": list(float(g["r_out"]) * np.sin(th))},
           "right": {"x": list(float(g["r_in"]) * np.cos(th)), "y": list(float(g["r_in"]) * np.sin(th))}}
    path = os.path.join(str(directory), "circle.json")
    with open(path, "w") as fh:
        json.dump(doc, fh)
    return path


def case_setup(name):
    """Decode a fixture name (tools/make_golden.py) into (track json, width, vehicle json, mode)."""
    import lap_time_optimization_b200 as ltk

    track = name.split("_")[0]
    width = 1.0 if "w100" in name else (0.6 if ("full" in name and track != "buckmore") else 0.8)
    veh = "MX5.json" if "mx5" in name else "tbr18.json"
    mode = "full" if "full" in name else "bayes"
    return ltk.data_path("tracks", track + ".json"), width, ltk.data_path("vehicles", veh), mode


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
        return cache[name]

    return load


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.abs(b)
