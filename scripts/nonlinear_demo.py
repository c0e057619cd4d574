"""The reference's --nonlinear stage end to end (trajectory_bayesian_nonlinear.py:229-270) on the B200 path:
random population scored in one pass, the 10 fastest refined by COBYLA (maxiter 2000) in lock step.
The reference on this class of host: 965 s for 100 candidates + 10 serial COBYLA runs."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lap_time_optimization_b200 as ltk

population = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
track = ltk.Track(ltk.data_path("tracks", "buckmore.json"), track_width=0.8, quiet=True)
traj = ltk.TrajectoryBayesianNonlinear(track, ltk.load_vehicle(ltk.data_path("vehicles", "tbr18.json")))
t0 = time.time()
laps, best, idx = traj.population_topk(traj.random_population_device(population, (2026, 1)), 10)
t_pop = time.time() - t0
took = traj.Nonlinear(population=population, starts=10, key=(2026, 1))
print(f"population of {population}: {t_pop * 1e3:.1f} ms (first call, includes context creation); best random lap {best[0]:.4f} s")
print(f"Nonlinear(population={population}, starts=10, maxiter=2000): {took:.1f} s on {os.cpu_count()} host cores; "
      f"best lap {traj.best_tau:.4f} s; COBYLA improved the best random candidate by {best[0] - traj.best_tau:.4f} s")
