"""Host-in / host-out latency of small batches (the optimiser loops of SURVEY section 8(f) N1):
`python scripts/latency_probe.py` prints microseconds per call for B = 1 .. 1024 through
`LapTimeEvaluator.lap_times` (ltk_eval_alphas_host: one CUDA graph per batch size; LTK_NO_GRAPH=1 launches
the same work directly), through the torch-staged route and, beside them, the kernels' own time (CUDA events)."""
import sys, time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lap_time_optimization_b200 import LapTimeEvaluator, Track, Vehicle, data_path  # noqa: E402

track = Track(json_path=data_path("tracks/buckmore.json"), track_width=1.0, quiet=True)
ev = LapTimeEvaluator(track, Vehicle(data_path("vehicles/tbr18.json"), quiet=True), mode="bayes")
torch = ev.torch
rng = np.random.default_rng(3)
for B in (1, 10, 44, 132, 1024, 8192):
    a = rng.uniform(0, 0.99, (B, ev.n_alpha))
    for _ in range(20):
        ev.lap_times(a)
    reps = 200
    t0 = time.perf_counter()
    for _ in range(reps):
        ev.lap_times(a)
    host_us = (time.perf_counter() - t0) / reps * 1e6
    for _ in range(20):
        ev._lap_times_staged(a)
    t0 = time.perf_counter()
    for _ in range(reps):
        ev._lap_times_staged(a)
    staged_us = (time.perf_counter() - t0) / reps * 1e6
    d_a = torch.from_numpy(a).to(ev.device)
    out = torch.empty(B, dtype=torch.float64, device=ev.device)
    kt = ev.kernel_times(d_a, out, reps=20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ev.lap_times_device(d_a, out)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B:5d}  lap_times (host in/out) {host_us:8.1f} us   torch-staged route {staged_us:8.1f} us   device pipeline {e0.elapsed_time(e1) / reps * 1e3:8.1f} us   "
          + "  ".join(f"{k} {v * 1e3:.1f}" for k, v in kt.items()))
