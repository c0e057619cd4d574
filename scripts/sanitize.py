"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / initcheck): every kernel once on small,
ragged batches of both vehicles and both K1 paths, the facades, top-k stages and the device generator."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lap_time_optimization_b200 as ltk

tj = ltk.data_path("tracks", "buckmore.json")
for veh, ns in (("tbr18", None), ("MX5", None), ("tbr18", 2501)):
    track = ltk.Track(tj, track_width=0.8, quiet=True)
    ev = ltk.LapTimeEvaluator(track, ltk.load_vehicle(ltk.data_path("vehicles", veh + ".json")), "bayes", ns, device=0)
    for B in (1, 33, 1000, 5000, 40000):
        a = ev.random_population_device(B, (1, B))
        lap = ev.lap_times_device(a)
        best, idx = ev.topk_device(lap, 10)
        g2, length = ev.curvature_objectives_device(a)
        # the selection fused into the sweep epilogue, its packed output and the gathered merge
        lap_f, best_f, idx_f, packed = ev.lap_times_topk_device(a, k=10, index_base=5, packed=True)
        packed2 = ev.lap_times_topk_device(a, k=10, index_base=10 ** 6, packed=True)[3]
        mb, mi = ev.merge_gathered_device(torch.cat([packed, packed2]), 2, 10, 10)
        torch.cuda.synchronize()
        assert torch.equal(lap_f, lap) and torch.equal(idx_f[:min(B, 10)] - 5, idx[:min(B, 10)])
        pairs = sorted((float(v), int(i)) for p_ in (packed, packed2)
                       for v, i in zip(p_[:10].view(torch.float64).tolist(), p_[10:].tolist()) if i >= 0)[:10]
        assert mi.tolist()[:len(pairs)] == [i for _, i in pairs] and mb.tolist()[:len(pairs)] == [v for v, _ in pairs]
        assert torch.isfinite(lap).all() and torch.isfinite(g2).all()
    ev.set_sweep_precision(32)
    lap32 = ev.lap_times_device(a)
    ev.set_sweep_precision(64)
    pr = ev.profile(a[0].cpu().numpy())
    assert abs(pr["lap"] - lap[0].item()) == 0.0
    ev.close()
traj = ltk.Trajectory(ltk.Track(tj, track_width=0.8, quiet=True), ltk.load_vehicle(ltk.data_path("vehicles", "tbr18.json")))
traj.update_velocity()
k = traj.path.curvature(traj.s[:-1])
print("sanitize pass done", traj.lap_time(), float(k.max()))
