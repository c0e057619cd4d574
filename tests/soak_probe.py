"""Soak: many device-generated populations against the C oracle, bit for bit (each 65,536-candidate population
exercises ~3e8 unguarded divisions and square roots).  Not collected by pytest; run on the GPU box:
    python tests/soak_probe.py [populations per vehicle] [tridiagonal|fitpack]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lap_time_optimization_b200 as ltk  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle.reference_port import OracleTrack, load_vehicle  # noqa: E402

npop = int(sys.argv[1]) if len(sys.argv) > 1 else 40
spline = sys.argv[2] if len(sys.argv) > 2 else "tridiagonal"
B = 65536
for trackname, veh in (("buckmore", "tbr18"), ("buckmore", "MX5"), ("whilton", "tbr18")):
    tj, vj = ltk.data_path("tracks", trackname + ".json"), ltk.data_path("vehicles", veh + ".json")
    ev = ltk.LapTimeEvaluator(ltk.Track(tj, track_width=0.8, quiet=True), ltk.load_vehicle(vj), "bayes", None, device=0,
                              spline=spline)
    co = c_oracle.COracle(OracleTrack(tj, 0.8), load_vehicle(vj), "bayes", None, device_sum_order=True, spline=spline)
    bad, t0 = 0, time.time()
    for i in range(npop):
        key = (777, i)
        got = ev.lap_times_device(ev.random_population_device(B, key)).cpu().numpy()
        a = np.random.Generator(np.random.Philox(key=np.array(key, dtype=np.uint64))).uniform(0.0, 0.99, (B, ev.n_alpha))
        want = co.lap_times(a)
        bad += int(np.sum(got != want))
    print(f"{trackname}/{veh} [{spline}]: {npop} populations x {B} candidates, {bad} lap times differ from the C oracle "
          f"({time.time() - t0:.0f} s)")
    ev.close()
