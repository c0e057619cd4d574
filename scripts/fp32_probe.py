"""fp32 sweep variant: error against the fp64 kernels and throughput (one B200)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lap_time_optimization_b200 as ltk
for veh in ("tbr18", "MX5"):
    track = ltk.Track(ltk.data_path("tracks", "buckmore.json"), track_width=0.8, quiet=True)
    ev = ltk.LapTimeEvaluator(track, ltk.load_vehicle(ltk.data_path("vehicles", veh + ".json")), "bayes", None, device=0)
    B = 65536
    pops = [ev.random_population_device(B, (5, i)) for i in range(8)]
    outs = [torch.empty(B, dtype=torch.float64, device="cuda") for _ in range(3)]
    res = {}
    for bits in (64, 32):
        ev.set_sweep_precision(bits)
        laps = ev.lap_times_device(pops[0]).cpu().numpy()
        res[bits] = laps
        for lanes in (1, 3):
            ev.run_resident((pops[i % 8] for i in range(12)), outs, 10, lanes=lanes)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ev.run_resident((pops[i % 8] for i in range(96)), outs, 10, lanes=lanes)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 96
            kt = ev.kernel_times(pops[0], outs[0], reps=5)
            bytes_per = 8 * 43 + (40 if bits == 64 else 28) * 846 + 8
            print(f"{veh} fp{bits} lanes {lanes}: {ms:.4f} ms/step {B / ms / 1e3:.1f} M evals/s, {B * bytes_per / ms / 1e6:.0f} GB/s algorithmic; kernels {kt}")
    rel = np.abs(res[32] - res[64]) / res[64]
    print(f"{veh}: fp32 sweeps vs fp64: median {np.median(rel):.2e} p99 {np.percentile(rel, 99):.2e} max {rel.max():.2e}")
    ev.close()
