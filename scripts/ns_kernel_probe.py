"""Developer probe (GPU box): per-kernel CUDA-event times at a given sampling density / population size.
    [LTK_K1=old] [LTK_K1_G=..] [LTK_K1_THREADS=..] python scripts/ns_kernel_probe.py NS B [fitpack]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import lap_time_optimization_b200 as ltk  # noqa: E402

ns, B = int(sys.argv[1]), int(sys.argv[2])
spline = sys.argv[3] if len(sys.argv) > 3 else "tridiagonal"
tj, vj = ltk.data_path("tracks", "buckmore.json"), ltk.data_path("vehicles", "tbr18.json")
ev = ltk.LapTimeEvaluator(ltk.Track(tj, track_width=0.8, quiet=True), ltk.load_vehicle(vj), "bayes", ns, device=0, spline=spline)
d_a = ev.random_population_device(B, (1, 2))
out = torch.empty(B, dtype=torch.float64, device="cuda")
ev.kernel_times(d_a, out, reps=1)
kt = ev.kernel_times(d_a, out, reps=3)
env = {k: v for k, v in os.environ.items() if k.startswith("LTK_")}
print(f"ns {ns} B {B} {spline} {env}: " + ", ".join(f"{k} {v:.3f} ms" for k, v in kt.items()) + f", sum {sum(kt.values()):.3f} ms", flush=True)
