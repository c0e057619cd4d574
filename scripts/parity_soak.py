"""Developer probe (GPU box): lap times of a large sample against the reference-equivalent port (live SciPy FITPACK /
numpy on all host cores), both spline modes: the distribution behind the 1e-9 claim.
    python scripts/parity_soak.py [rows=65536] [vehicle=tbr18]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lap_time_optimization_b200 as ltk  # noqa: E402
from oracle.reference_port import lap_times_pool  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
veh = sys.argv[2] if len(sys.argv) > 2 else "tbr18"
tj, vj = ltk.data_path("tracks", "buckmore.json"), ltk.data_path("vehicles", veh + ".json")
a = np.random.default_rng(2026).uniform(0.0, 0.99, (rows, 43))
ref = lap_times_pool(tj, 0.8, vj, a, "bayes")  # before CUDA is initialised: the pool forks
out = {"rows": rows, "vehicle": veh, "host_cores": os.cpu_count()}
for mode in ("tridiagonal", "fitpack"):
    ev = ltk.LapTimeEvaluator(ltk.Track(tj, track_width=0.8, quiet=True), ltk.load_vehicle(vj), "bayes", None, device=0, spline=mode)
    rel = np.abs(ev.lap_times(a) - ref) / ref
    out[mode] = {"median": float(np.median(rel)), "p99": float(np.percentile(rel, 99)), "p99.9": float(np.percentile(rel, 99.9)),
                 "max": float(rel.max()), "count_over_1e-9": int((rel > 1e-9).sum()), "count_over_1e-10": int((rel > 1e-10).sum()),
                 "bit_equal": int((rel == 0).sum())}
    ev.close()
print(json.dumps(out))
