"""Developer probe (GPU box): lap times of a large sample against the reference-equivalent port (live SciPy FITPACK /
numpy on all host cores), both spline modes, against BOTH arithmetics the unmodified reference has on x86 hosts:
numpy's default dispatch on this host (AVX512: `x ** 1.5` through its vendored SVML) and numpy's baseline dispatch
(NPY_DISABLE_CPU_FEATURES: libm `pow`) -- the distribution behind the 1e-9 claim.
    python scripts/parity_soak.py [rows=65536] [vehicle=tbr18]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def port_laps(a, tj, vj, baseline):
    from oracle.reference_port import lap_times_baseline_dispatch, lap_times_pool

    if baseline:
        return lap_times_baseline_dispatch(tj, 0.8, vj, a, "bayes"), "False"
    return lap_times_pool(tj, 0.8, vj, a, "bayes"), "host default"


def main():
    import lap_time_optimization_b200 as ltk

    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    veh = sys.argv[2] if len(sys.argv) > 2 else "tbr18"
    tj, vj = ltk.data_path("tracks", "buckmore.json"), ltk.data_path("vehicles", veh + ".json")
    a = np.random.default_rng(2026).uniform(0.0, 0.99, (rows, 43))
    refs = {}
    for name, baseline in (("host_dispatch", False), ("baseline_dispatch", True)):
        laps, avx = port_laps(a, tj, vj, baseline)
        refs[name] = laps
        print(f"port, {name}: AVX512 pow in use: {avx}", file=sys.stderr)
    rel_refs = np.abs(refs["host_dispatch"] - refs["baseline_dispatch"]) / refs["baseline_dispatch"]
    out = {"rows": rows, "vehicle": veh, "host_cores": os.cpu_count(),
           "reference_against_itself": {"what": "the port under numpy's host dispatch against the port under numpy's baseline dispatch",
                                        "median": float(np.median(rel_refs)), "p99": float(np.percentile(rel_refs, 99)),
                                        "max": float(rel_refs.max()), "count_over_1e-9": int((rel_refs > 1e-9).sum()),
                                        "bit_equal": int((rel_refs == 0).sum())}}
    for mode in ("tridiagonal", "fitpack"):
        ev = ltk.LapTimeEvaluator(ltk.Track(tj, track_width=0.8, quiet=True), ltk.load_vehicle(vj), "bayes", None, device=0, spline=mode)
        got = ev.lap_times(a)
        ev.close()
        out[mode] = {}
        for name, ref in refs.items():
            rel = np.abs(got - ref) / ref
            out[mode][name] = {"median": float(np.median(rel)), "p99": float(np.percentile(rel, 99)),
                               "p99.9": float(np.percentile(rel, 99.9)), "max": float(rel.max()),
                               "count_over_1e-9": int((rel > 1e-9).sum()), "count_over_1e-10": int((rel > 1e-10).sum()),
                               "bit_equal": int((rel == 0).sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
