"""Do two populations in flight on two streams overlap usefully?  (K1 of one with K23 of the other.)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lap_time_optimization_b200 as ltk

dev = torch.device("cuda", 0)
tj, vj = ltk.data_path("tracks", "buckmore.json"), ltk.data_path("vehicles", "tbr18.json")
track = ltk.Track(tj, track_width=0.8, quiet=True)
veh = ltk.load_vehicle(vj)
B = int(os.environ.get("B", 65536))
NS = int(os.environ.get("NS", 4))
evs = [ltk.LapTimeEvaluator(track, veh, "bayes", None, device=0, max_workspace_bytes=4 << 30) for _ in range(NS)]
pops = [torch.as_tensor(np.random.default_rng(i).uniform(0, 0.99, (B, 43))).to(dev) for i in range(4)]
laps = [torch.empty(B, dtype=torch.float64, device=dev) for _ in range(NS)]
streams = [torch.cuda.Stream(dev) for _ in range(NS)]
K = 20


def run(ns):
    def fn():
        for i in range(12 * K // 12):
            s = i % ns
            with torch.cuda.stream(streams[s]):
                evs[s].lap_times_device(pops[i % 4], out=laps[s])
                evs[s].topk_device(laps[s], 10)
    return fn


K = 48
for ns in range(1, NS + 1):
    fn = run(ns)
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{ns} stream(s): {1e3 * dt / K:.4f} ms per population  ({B * K / dt / 1e6:.1f} M evals/s)")
