"""Where a K1b CTA spends its life: phase time stamps (%globaltimer, thread 0) of every CTA of one launch.
Needs a developer build:  make -C lap_time_optimization_b200/csrc variant NAME=clock EXTRA=-DLTK_K1B_CLOCK
    LTK_LIB_PATH=.../variants/libltk_clock.so python scripts/k1b_phase_probe.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lap_time_optimization_b200 as ltk
from lap_time_optimization_b200 import _native

track = ltk.Track(ltk.data_path("tracks", "buckmore.json"), track_width=0.8, quiet=True)
ev = ltk.LapTimeEvaluator(track, ltk.load_vehicle(ltk.data_path("vehicles", "tbr18.json")), "bayes", None, device=0)
B = 65536
pop = ev.random_population_device(B, (3, 1))
out = torch.empty(B, dtype=torch.float64, device="cuda")
for _ in range(3):
    ev.lap_times_device(pop, out=out)
torch.cuda.synchronize()
lib = _native.load()
n_cta, slots = B // 4, 8
buf = np.zeros(n_cta * slots, dtype=np.int64)
lib.ltk_debug_k1b_clocks.restype = C.c_int
assert lib.ltk_debug_k1b_clocks(buf.ctypes.data_as(C.c_void_p), C.c_longlong(buf.size)) == 0
t = buf.reshape(n_cta, slots)[:, :6].astype(np.float64)
t0 = t[:, 0].min()
names = ["L  (bulk copy + control points)", "C  (interval records)", "K  (sample loop)", "A  (arg-max, barrier)", "W  (rotated write-out)"]
print(f"kernel span {(t[:, 5].max() - t0) / 1e3:.1f} us, {n_cta} CTAs, mean CTA life {(t[:, 5] - t[:, 0]).mean() / 1e3:.2f} us")
for i, nm in enumerate(names):
    d = t[:, i + 1] - t[:, i]
    print(f"  {nm:34s} mean {d.mean() / 1e3:6.2f} us  p10 {np.percentile(d, 10) / 1e3:6.2f}  p90 {np.percentile(d, 90) / 1e3:6.2f}")
# CTAs resident at a time
starts, ends = np.sort(t[:, 0]), np.sort(t[:, 5])
mid = t0 + 0.5 * (t[:, 5].max() - t0)
print("  resident CTAs at mid-kernel:", int((starts <= mid).sum() - (ends <= mid).sum()), "(148 SMs)")
