"""BASELINE.json configs 3-5 on ONE B200 (the 8-GPU versions shard the same populations by rows):
throughput with resident inputs, and parity of a random subsample against the C oracle (bit for bit).

    python tests/configs_probe.py [--quick]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
import lap_time_optimization_b200 as ltk  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle.reference_port import OracleTrack, load_vehicle  # noqa: E402

quick = "--quick" in sys.argv
dev = torch.device("cuda", 0)
tj = ltk.data_path("tracks", "buckmore.json")
CASES = [  # name, vehicle, candidates on this GPU, ns
    ("config 3: MX5, 2^20 candidates (one GPU holds what eight would share)", "MX5", 1 << 20, None),
    ("config 4: TBR18 database, 2^20 device-generated candidates", "tbr18", 1 << 20, None),
    ("config 5: TBR18, ns = 10001, 524,288 candidates (= 4M / 8 GPUs)", "tbr18", 1 << 19, 10001),
]
for name, veh, B, ns in CASES:
    if quick:
        B //= 16
    vj = ltk.data_path("vehicles", veh + ".json")
    track = ltk.Track(tj, track_width=0.8, quiet=True)
    ev = ltk.LapTimeEvaluator(track, ltk.load_vehicle(vj), "bayes", ns, device=0)
    key = (2026, 1018)
    d_a = ev.random_population_device(B, key)
    d_lap = torch.empty(B, dtype=torch.float64, device=dev)
    ev.lap_times_device(d_a, out=d_lap)  # warm-up on the same path (creates lanes / workspaces)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        ev.lap_times_device(d_a, out=d_lap)
        best, idx = ev.topk_device(d_lap, 10)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = ev.ns - 1
    bytes_per = 8 * ev.n_alpha + 40 * n + 8
    print(f"{name}\n   {ms:9.2f} ms per pass, {B / ms / 1e3:8.2f} M evals/s, {B * bytes_per / ms / 1e6:7.1f} GB/s algorithmic "
          f"({B * bytes_per / ms / 1e6 / 6545.6:.1%} of the HBM peak), chunk {ev.max_batch()} candidates")
    # parity: a random subsample against the C oracle, and the top-10 against a host sort of the device laps
    sub = np.random.default_rng(1).choice(B, size=256 if ns else 2048, replace=False)
    a_host = d_a[torch.as_tensor(sub, device=dev)].cpu().numpy()
    co = c_oracle.COracle(OracleTrack(tj, 0.8), load_vehicle(vj), "bayes", ns, device_sum_order=True)
    t0 = time.perf_counter()
    want = co.lap_times(a_host)
    dt = time.perf_counter() - t0
    got = d_lap.cpu().numpy()
    ok = np.array_equal(got[sub], want)
    order = np.lexsort((np.arange(B), got))[:10]
    print(f"   subsample of {len(sub)} bit-identical to the C oracle: {ok} (C oracle: {len(sub) / dt:.0f} evals/s on all host threads); "
          f"top-10 indices equal a stable host sort: {np.array_equal(order, idx.cpu().numpy())}; best lap {best[0].item():.4f} s")
    ev.close()
    del d_a, d_lap
    torch.cuda.empty_cache()
