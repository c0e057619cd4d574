import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import lap_time_optimization_b200 as ltk
for veh in ("tbr18", "MX5"):
    ev = ltk.LapTimeEvaluator(ltk.Track(ltk.data_path("tracks","buckmore.json"), track_width=0.8, quiet=True), ltk.load_vehicle(ltk.data_path("vehicles", veh+".json")), "bayes", None, device=0)
    worst, med = 0.0, []
    for i in range(20):
        a = ev.random_population_device(65536, (55, i))
        ev.set_sweep_precision(64); l64 = ev.lap_times_device(a).clone()
        ev.set_sweep_precision(32); l32 = ev.lap_times_device(a)
        rel = ((l32 - l64).abs() / l64)
        worst = max(worst, rel.max().item()); med.append(rel.median().item())
    print(f"{veh}: fp32 sweeps vs fp64 over 1,310,720 candidates: median {np.median(med):.2e}, max {worst:.2e}")
    ev.close()
