"""Time line of the pipeline kernels with three populations in flight (ltk_trace_*): who overlaps whom."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lap_time_optimization_b200 as ltk

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
track = ltk.Track(ltk.data_path("tracks", "buckmore.json"), track_width=0.8, quiet=True)
ev = ltk.LapTimeEvaluator(track, ltk.load_vehicle(ltk.data_path("vehicles", "tbr18.json")), "bayes", None, device=0)
B = 65536
pops = [ev.random_population_device(B, (3, i)) for i in range(8)]
outs = [torch.empty(B, dtype=torch.float64, device="cuda") for _ in range(lanes)]
ev.run_resident((pops[i % 8] for i in range(30)), outs, 10, lanes=lanes)
torch.cuda.synchronize()
ev.trace_begin(64)
ev.run_resident((pops[i % 8] for i in range(24)), outs, 10, lanes=lanes)
torch.cuda.synchronize()
rows = ev.trace_read()
t_end = max(r[3] for r in rows)
print(f"{lanes} lanes, 24 populations: {t_end:.3f} ms total, {t_end / 24:.4f} ms per population")
for lane, kern, t0, t1 in rows[9 * 3:9 * 3 + 27]:  # a steady-state window
    print(f"  lane {lane} {kern:4s} {t0:8.3f} -> {t1:8.3f}  ({t1 - t0:.3f} ms)" + "   " * lane + " " + "#" * max(1, int((t1 - t0) * 40)))
import collections
dur = collections.defaultdict(list)
for lane, kern, t0, t1 in rows[9:]:
    dur[kern].append(t1 - t0)
print({k: round(sum(v) / len(v), 4) for k, v in dur.items()}, "mean event-to-event duration per kernel while overlapped")
