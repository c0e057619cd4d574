#!/usr/bin/env python3
"""bench.py -- lap-time evaluations/sec on Buckmore + TBR18 (BASELINE.json), B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" scores one population of `--candidates` (default 65,536 = BASELINE.json configs[1]) random
alpha vectors per GPU: K1a spline solve -> K1b curvature (rotated write-out) -> K23 forward + backward
sweeps with the lap sum and the top-10 selected in the sweep's epilogue -- three kernel launches, one
ltk_eval_alphas_topk call (-> one all-gather of 160 B/rank + merge when N > 1).  Candidates are
independent, so ranks get disjoint populations and the scaling is weak.  `--lanes` (default 3) steps are
in flight at a time, each on its own stream (LapTimeEvaluator.lanes).  Prints ONE JSON line (rank 0).

`value`   : whole-job evaluations/s, inputs already resident in HBM, CUDA events around exactly K steps,
            max over ranks.
`e2e`     : same metric through the public host API (LapTimeEvaluator.stream_populations): every step's
            population goes pinned host memory -> H2D -> pipeline -> top-k -> D2H of all lap times and
            the top-10 inside the timed region; copies and kernels of different steps overlap.
`roofline`: the dominant kernel's algorithmic bytes / its CUDA-event duration (one population at a time,
            ltk_eval_alphas_timed) against the measured HBM peak (MEASURED_PEAKS.json, burst figure;
            fallback 6650 GB/s per B200_PROFILING.md); `pipeline` = the same for the whole step.
`cpu_baseline` / `--impl reference`: the reference's own CPU path (oracle/reference_port.py: the
            reference restated one candidate at a time with the same SciPy/numpy calls, pinned
            bit-for-bit to the unmodified reference by tests/golden) on all host cores (warm fork pool).
`parity`  : (N = 1) lap times of 4,096 rows of the timed population against that port, for both spline modes
            (median / p99 / max / count over 1e-9) and whether the top-10 of the sample is identical.
`variants`: short runs of the same workload in the FITPACK spline mode (the reference's own bits) and with
            the optional fp32 sweeps.
`topk_identical`: (N > 1) every rank's lap times of one step gathered to rank 0, stable host sort, compared with
            the all-gathered + merged device top-10 (trajectory_bayesian_nonlinear.py:253-257 across ranks).
`configs` : short runs of BASELINE.json configs 3, 4, 5 at their named total sizes, strong-sharded over the N
            ranks (skip with --no-configs): evals/s, HBM fraction, and a bit-exact check of a subsample against
            the C oracle.
`h2d_probe`: pinned host -> device copy bandwidth of this rank while all N ranks copy at once.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRACK, VEHICLE, WIDTH = "buckmore", "tbr18", 0.8
TOPK = 10
METRIC = "lap_time_evals_per_sec"
UNIT = "evals/s"
N_INPUT_SETS = 8  # distinct resident populations cycled through the steps (8 x 22.5 MB > 126 MB L2)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=500)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--candidates", type=int, default=65536, help="candidates per GPU per step")
    p.add_argument("--vehicle", default=VEHICLE, choices=["tbr18", "MX5"])
    p.add_argument("--ns", type=int, default=None, help="samples per lap incl. end point (default ceil(track length) = 847)")
    p.add_argument("--cpu-sample", type=int, default=4096, help="candidates scored by the CPU baseline / parity sample")
    p.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 3/4/5 block")
    p.add_argument("--no-variants", action="store_true", help="skip the FITPACK-mode and fp32 variant runs")
    p.add_argument("--spline", default="tridiagonal", choices=["tridiagonal", "fitpack"],
                   help="spline arithmetic of the timed run (fitpack = SciPy FITPACK's own operation order)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--sweep-bits", type=int, default=64, choices=[64, 32],
                   help="32 = the optional fp32 variant of the velocity sweeps (spline and curvature stay fp64)")
    p.add_argument("--lanes", type=int, default=3, help="populations in flight per GPU (1 = strictly one step at a time)")
    p.add_argument("--e2e-slots", type=int, default=None,
                   help="input/result buffer sets of the end-to-end pipeline (default 2 x lanes)")
    return p.parse_args()


def data_paths(vehicle):
    import lap_time_optimization_b200 as ltk

    return ltk.data_path("tracks", TRACK + ".json"), ltk.data_path("vehicles", vehicle + ".json")


def workload_config(args, n_alpha, ns):
    return {"workload": f"{TRACK} track + {args.vehicle} vehicle, width {WIDTH}, random alpha ~ U[0,0.99)^{n_alpha}, "
                        f"{args.candidates} candidates per GPU per step, {ns - 1} samples per lap, top-{TOPK}",
            "candidates_per_gpu": args.candidates, "n_alpha": n_alpha, "samples_per_lap": ns - 1, "topk": TOPK,
            "steps_in_flight": getattr(args, "lanes", 1),
            "l2": f"{N_INPUT_SETS} distinct resident populations cycled (inputs {N_INPUT_SETS}x{args.candidates * n_alpha * 8 / 1e6:.1f} MB) and "
                  f"{2 * (ns - 1) * args.candidates * 8 / 1e9:.2f} GB of staged intermediates rewritten every step (> 126 MB L2)"}


# ------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------
def cpu_rate(args, sample, processes):
    """evals/s of the reference-equivalent CPU port on `processes` host processes (pool started and warmed
    before the clock, like the reference arm)."""
    from oracle.reference_port import LapPool

    tj, vj = data_paths(args.vehicle)
    n_alpha = 43
    a = np.random.default_rng(1002).uniform(0.0, 0.99, (sample, n_alpha))
    pool = LapPool(tj, WIDTH, vj, "bayes", args.ns, processes)
    try:
        pool.lap_times(np.random.default_rng(1).uniform(0.0, 0.99, (4 * processes, n_alpha)))
        t0 = time.perf_counter()
        laps = pool.lap_times(a)
        dt = time.perf_counter() - t0
    finally:
        pool.close()
    return sample / dt, laps, a


def port_laps_baseline_dispatch(args, a):
    """The reference-equivalent port on the same rows with numpy's AVX512 dispatch switched off (a process of its own).
    The unmodified reference evaluates `(dx**2 + dy**2) ** 1.5` (path.py:58) with numpy's vendored SVML `pow` on AVX512
    hosts and with libm's elsewhere, and its TBR18 lap times move by up to 6.5e-9 between the two
    (profiles/README.md): parity is reported against both."""
    from oracle.reference_port import lap_times_baseline_dispatch

    tj, vj = data_paths(args.vehicle)
    return lap_times_baseline_dispatch(tj, WIDTH, vj, a, "bayes", args.ns)


def rel_stats(ours, ref):
    rel = np.abs(np.asarray(ours) - np.asarray(ref)) / np.abs(ref)
    return {"n": int(rel.size), "median": float(np.median(rel)), "p99": float(np.percentile(rel, 99)),
            "max": float(rel.max()), "count_over_1e-9": int((rel > 1e-9).sum())}


def stable_topk(laps, k):
    order = np.argsort(np.asarray(laps), kind="stable")[:k]  # sorted(...)[0:k], ties keep population order
    return order.astype(np.int64), np.asarray(laps)[order]


def run_reference(args):
    """The reference arm: the reference's CPU path (reference-equivalent port, all host cores) on bounded
    samples of the same workload -- one sample of `sample` candidates per step, sized so that the whole
    --steps/--warmup run stays within about two minutes."""
    from oracle.reference_port import LapPool

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total_steps = args.warmup + args.steps
    budget = 110.0 * 230.0 * cores  # candidates affordable in ~110 s at ~230 evals/s/core (measured 234-255)
    sample = int(max(2 * cores, min(args.cpu_sample, 64 * cores, budget / max(total_steps, 1))))
    tj, vj = data_paths(args.vehicle)
    pool = LapPool(tj, WIDTH, vj, "bayes", args.ns, cores)
    rates = []
    for i in range(total_steps):
        a = np.random.default_rng(1002 + i).uniform(0.0, 0.99, (sample, 43))
        t0 = time.perf_counter()
        pool.lap_times(a)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            rates.append(sample / dt)
    pool.close()
    value = float(np.mean(rates))
    ns = args.ns or 847
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 43, ns),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} candidates per step, fork pool of {cores} processes, "
                                       "oracle/reference_port.py (reference restated with the same SciPy FITPACK / numpy calls)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 2 ms from a thread (the timed region
    of a 20-step run is 15 ms: nvidia-smi's 20 ms loop gave it one sample); nvidia-smi -lms as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.nvml, self.stop_flag = [], None, index, None, False

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.index]) if self.index < len(ids) else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": int(getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                "hw_thermal_slowdown": int(getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                "sw_thermal_slowdown": int(getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                "sw_power_cap": int(getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4))}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(get_reasons(self.handle))
                self.rows.append((time.perf_counter(), mhz, [k for k, b in bits.items() if mask & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.nvml is not None:
            time.sleep(0.01)
            self.stop_flag = True
            picked = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
            sm = [r[1] for r in picked]
            reasons = sorted({x for r in picked for x in r[2]})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(picked), "source": "nvml, 2 ms polling"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        picked = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows[-3:]]
        for r in picked:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(picked), "source": "nvidia-smi -lms 20"}


def fp64_pipe_model(args, B, n, ms_per_step, clocks):
    """FP64-pipe occupancy of one step under the instruction-cost model of profiles/README.md.
    Cycles per sample and scheduler: K1b 75 (33 FP64 instructions), K23 forward 87 + backward 70 + lap term 23
    (fp64 sweeps; the fp32 variant leaves only K1b and the fp64 lap sum on this pipe)."""
    per_sample = 75 + (87 + 70 + 23 if args.sweep_bits == 64 else 4)
    sms, schedulers = 148, 4
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    need = B * n * per_sample / 32.0 / (sms * schedulers)
    have = ms_per_step * 1e-3 * mhz * 1e6
    return {"model_cycles_per_sample": per_sample, "frac": need / have, "sm_mhz": mhz,
            "peak": "64 DFMA/clk/SM measured (tools/ubench/fp64_lat.cu): 37.2 TFLOP/s at 1965 MHz"}


def issue_model(args, B, n, ms_per_step, clocks):
    """Issue-slot occupancy of one step under the cost model of DESIGN.md section 3 ("What binds"): one cycle per
    instruction, two per FP64 instruction, three per DFMA with three distinct register operands (tools/ubench), applied
    to the SASS instruction counts of the kernels' loops (`python tools/issue_model.py lap_time_optimization_b200/libltk.so
    <kernel>` prints them).  Cycles per sample and warp: K23 277 (two row pairs: phase 1 = a 261-instruction block with 160
    FP64 and 12 three-operand DFMAs + the ~60-cycle loop header = 493 cycles, phase 2 = 315 / 200 / 20 + ~72 = 607; mean 550),
    K1b 135 (sample loop 90 = 50 instructions, 30 FP64, 10 three-operand DFMAs; interval switches 7, write-out 10, per-CTA
    phases 28), K1a 20.  fp64 sweeps only; this is the builder's model, not a counter."""
    if args.sweep_bits != 64 or args.spline != "tridiagonal":
        return None
    per_sample = 277 + 135 + 20
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    need = B * n * per_sample / 32.0 / (148 * 4)
    have = ms_per_step * 1e-3 * mhz * 1e6
    return {"model_cycles_per_sample_and_warp": per_sample, "frac": need / have, "sm_mhz": mhz}


def _numpy_simd():
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
        return {k for k, v in feats.items() if v}
    except Exception:
        return set()


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# extra blocks of the GPU arm: rank-0 identity checks, variants, BASELINE configs 3/4/5, H2D probe
# ------------------------------------------------------------------------------------------------
class Dist:
    """The little this file needs from torch.distributed, also valid at world size 1."""

    def __init__(self, torch, dist, world, rank, dev):
        self.torch, self.dist, self.world, self.rank, self.dev = torch, dist, world, rank, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_ms(self, ms):
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_rows(self, t):
        """Concatenate equally sized 1-D CUDA tensors of all ranks (rank order); every rank gets the result."""
        if self.world == 1:
            return t
        out = self.torch.empty(self.world * t.numel(), dtype=t.dtype, device=self.dev)
        self.dist.all_gather_into_tensor(out, t.contiguous())
        return out

    def all_true(self, flag):
        t = self.torch.tensor([1 if flag else 0], dtype=self.torch.int32, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


def topk_identity(D, ev, pop, base, finish):
    """One population per rank through the production path (pipeline -> local top-k -> all-gather -> merge);
    then EVERY lap time goes to rank 0, which sorts them on the host (stable: what `sorted(...)[0:10]` does at
    trajectory_bayesian_nonlinear.py:253-257) and compares indices and values with the merged device result."""
    d_lap, best, idx, packed = ev.lap_times_topk_device(pop, k=TOPK, index_base=base, packed=True)
    if finish is not None:
        best, idx = finish(packed) if getattr(finish, "packed", False) else finish(best, idx)
    all_laps = D.gather_rows(d_lap).cpu().numpy()  # rank r's rows sit at [r * B, (r + 1) * B) = its global indices
    h_idx, h_best = stable_topk(all_laps, TOPK)
    same = bool(np.array_equal(h_idx, idx.cpu().numpy()) and np.array_equal(h_best, best.cpu().numpy()))
    return D.all_true(same)


def short_run(D, ev, dev_sets, d_laps, steps, lanes, base, finish):
    """evals/s of `steps` resident populations (same protocol as the main timed region, shorter)."""
    torch = D.torch
    B = dev_sets[0].shape[0]
    run = lambda n: ev.run_resident((dev_sets[i % len(dev_sets)] for i in range(n)), d_laps, TOPK, index_base=base,  # noqa: E731
                                    finish=finish, lanes=lanes)
    run(3)
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps)
    e1.record()
    D.barrier()
    ms = D.max_ms(e0.elapsed_time(e1))
    return {"value": D.world * B * steps / (ms * 1e-3), "unit": UNIT, "steps": steps, "ms_per_step": ms / steps}


def c_oracle_for(vehicle, ns, spline="tridiagonal"):
    from oracle.c_oracle import COracle
    from oracle.reference_port import OracleTrack, load_vehicle

    tj, vj = data_paths(vehicle)
    return COracle(OracleTrack(tj, WIDTH), load_vehicle(vj), "bayes", ns, device_sum_order=True, spline=spline)


def run_config(D, ltk, local, tag, vehicle, ns, total, key, timed_generation, peak, reps):
    """One BASELINE.json config at its named TOTAL size, strong-sharded: rank r generates rows
    [r * total / N, (r + 1) * total / N) of ONE Philox population on its GPU, scores them, takes its local
    top-10 with global indices; all-gather + merge.  Checks: merged top-10 == stable host sort of all lap
    times; a subsample of rank 0's shard == the C oracle bit for bit; the device population == numpy's Philox."""
    from lap_time_optimization_b200.distributed import PackedTopkGather, shard_bounds

    torch = D.torch
    tj, vj = data_paths(vehicle)
    ev = ltk.LapTimeEvaluator(ltk.Track(tj, track_width=WIDTH, quiet=True), ltk.load_vehicle(vj), "bayes", ns, device=local)
    lo, hi = shard_bounds(total, D.rank, D.world)
    Bl = hi - lo
    out = torch.empty(Bl, dtype=torch.float64, device=D.dev)
    d_a = ev.random_population_device(Bl, key, first_row=lo)

    def once(generate):
        a = ev.random_population_device(Bl, key, first_row=lo) if generate else d_a
        d_lap, best, idx, packed = ev.lap_times_topk_device(a, out=out, k=TOPK, index_base=lo, packed=True)
        if D.world > 1:
            best, idx = PackedTopkGather(ev, TOPK)(packed)
        return a, d_lap, best, idx

    # warm-up on a slice: lanes, workspaces and the top-k scratch exist before the clock starts
    ev.lap_times_device(d_a[:min(Bl, 3 * ev.WAVE)], out=out[:min(Bl, 3 * ev.WAVE)])
    ev.topk_device(out, TOPK, index_base=lo)
    once(False)  # one full pass (cross-rank gather and merge included) before the clock starts
    if timed_generation:  # two generated populations are alive at a time: let the caching allocator own both blocks
        keep = once(True)
        once(True)
        del keep
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        a, d_lap, best, idx = once(timed_generation)
    e1.record()
    D.barrier()
    ms = D.max_ms(e0.elapsed_time(e1)) / reps
    n = ev.ns - 1
    a_staged = 8 * ev.n_alpha + 40 * n + 8
    # identity of the global top-10
    if total % D.world == 0:
        all_laps = D.gather_rows(d_lap).cpu().numpy()
        h_idx, h_best = stable_topk(all_laps, TOPK)
        same = D.all_true(bool(np.array_equal(h_idx, idx.cpu().numpy()) and np.array_equal(h_best, best.cpu().numpy())))
    else:
        same = None
    res = {"workload": tag, "vehicle": vehicle, "candidates_total": total, "candidates_per_gpu": Bl,
           "samples_per_lap": n, "scaling": "strong", "ms": ms, "value": total / (ms * 1e-3), "unit": UNIT,
           "hbm_frac": a_staged * Bl / (ms * 1e-3) / 1e9 / peak, "bytes_per_candidate": a_staged,
           "population": "device-generated (ltk_random_uniform: numpy Philox stream)" +
                         (", generation inside the timed region" if timed_generation else ", resident"),
           "topk_identical": same}
    if D.rank == 0:
        rows = np.sort(np.random.default_rng(11).choice(Bl, min(Bl, 2048 if n < 2000 else 128), replace=False))
        h_a = a[torch.as_tensor(rows, device=D.dev)].cpu().numpy()
        want = c_oracle_for(vehicle, ev.ns).lap_times(h_a)
        res["subsample_bit_exact"] = {"rows": int(rows.size), "equal": bool(np.array_equal(want, d_lap.cpu().numpy()[rows]))}
        m = min(Bl, 1024)
        h_pop = np.random.Generator(np.random.Philox(key=list(key))).uniform(0.0, 0.99, (m, ev.n_alpha))
        res["population_equals_numpy_philox"] = bool(np.array_equal(h_pop, a[:m].cpu().numpy()))
    ev.close()
    del out, d_a, a, d_lap
    torch.cuda.empty_cache()
    return res


def h2d_probe(D, nbytes, reps=96, warm=32):
    """Pinned host -> device bandwidth of every rank while ALL ranks copy at the same time (the e2e path's
    upload of one population, back to back).  The link needs ~0.5 GB of traffic to reach its steady rate (measured
    on the pool's boxes: 18 GB/s over the first 24 copies of 22.5 MB, 24-36 GB/s over 100), hence the long warm-up."""
    torch = D.torch
    h = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
    d = torch.empty(nbytes // 8, dtype=torch.float64, device=D.dev)
    st = torch.cuda.Stream(D.dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        for _ in range(warm):
            d.copy_(h, non_blocking=True)
    D.barrier()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        e1.record(st)
    D.barrier()
    gbs = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=D.dev)
    allv = D.gather_rows(t).cpu().numpy()
    return {"bytes_per_copy": int(nbytes), "copies": reps, "concurrent_ranks": D.world,
            "gbs_per_rank": [round(float(x), 2) for x in allv], "gbs_total": round(float(allv.sum()), 1)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import lap_time_optimization_b200 as ltk
    from lap_time_optimization_b200 import _native
    from lap_time_optimization_b200.distributed import PackedTopkGather, allgather_topk

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_result = None
    cpu_laps_base = None
    cpu_one_core = None
    if world == 1 and not args.no_cpu_baseline:
        # the CPU baseline forks a worker pool: do it before this process owns a CUDA context
        cpu_result = cpu_rate(args, args.cpu_sample, os.cpu_count() or 1)
        cpu_one_core = cpu_rate(args, 192, 1)[0]  # the rate at P cores and at one core
        try:
            cpu_laps_base = port_laps_baseline_dispatch(args, cpu_result[2])
        except Exception:  # the probe is extra evidence, not part of the contract
            cpu_laps_base = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    tj, vj = data_paths(args.vehicle)
    track = ltk.Track(tj, track_width=WIDTH, quiet=True)
    ev = ltk.LapTimeEvaluator(track, ltk.load_vehicle(vj), "bayes", args.ns, device=local, spline=args.spline)
    if args.sweep_bits == 32:
        ev.set_sweep_precision(32)
    B, na, ns = args.candidates, ev.n_alpha, ev.ns
    base = rank * B  # global index of this rank's first candidate
    # resident inputs: N_INPUT_SETS distinct populations per rank
    host_sets = [np.random.default_rng(1002 + 7919 * rank + i).uniform(0.0, 0.99, (B, na)) for i in range(N_INPUT_SETS)]
    dev_sets = [torch.as_tensor(h).to(dev) for h in host_sets]
    LANES = args.lanes  # populations in flight (LapTimeEvaluator.lanes): kernels of different steps overlap
    d_laps = [torch.empty(B, dtype=torch.float64, device=dev) for _ in range(LANES)]
    d_lap = d_laps[0]
    # the cross-rank step of every population: one all-gather of the packed top-10 list + one merge launch
    # (LTK_BENCH_FINISH=unpacked: the five-launch route through allgather_topk, for A/B)
    if world > 1 and os.environ.get("LTK_BENCH_FINISH") == "unpacked":
        finish = lambda b, ix: allgather_topk(b, ix, TOPK, merge=ev.merge_topk_device)
    else:
        finish = PackedTopkGather(ev, TOPK) if world > 1 else None

    def run_steps(nsteps):
        return ev.run_resident((dev_sets[i % N_INPUT_SETS] for i in range(nsteps)), d_laps, TOPK, index_base=base,
                               finish=finish, lanes=LANES)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    run_steps(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _native.launch_count()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    best, idx = run_steps(args.steps)
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = _native.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public host API: pinned host populations -> H2D -> pipeline -> top-k
    #      (-> all-gather + merge) -> D2H of every lap time and the top-k.  `stream_populations` double-
    #      buffers: the H2D copy of step i+1 and the D2H of step i-1 overlap the kernels of step i; every
    #      byte of every step still crosses PCIe inside the timed region. ------------------------------
    pin_in = [torch.as_tensor(h).pin_memory() for h in host_sets[:4]]
    e2e_finish = None if os.environ.get("LTK_BENCH_E2E_NO_FINISH") else finish  # diagnostic: e2e without the all-gather
    if world > 1 and os.environ.get("LTK_BENCH_E2E_FINISH") == "merge_only":    # diagnostics: which half of it costs
        e2e_finish = lambda b, ix: ev.merge_topk_device(torch.cat([b] * world), torch.cat([ix] * world), TOPK)
    if world > 1 and os.environ.get("LTK_BENCH_E2E_FINISH") == "nccl_only":
        def e2e_finish(b, ix):
            packed = torch.cat([b.view(torch.int64), ix])
            gathered = torch.empty(world * packed.numel(), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(gathered, packed)
            return b, ix
    def e2e_run(nsteps):
        checksum = 0.0
        for laps, best_h, idx_h in ev.stream_populations((pin_in[i % 4] for i in range(nsteps)), TOPK,
                                                         index_base=base, index_stride=0, finish=e2e_finish,
                                                         lanes=LANES, slots=args.e2e_slots):
            checksum += float(best_h[0]) + float(laps[-1])  # results are consumed on the host
        return checksum

    # every buffer set exists before the timed region, and the host link is at its steady rate (it needs ~0.5 GB of
    # traffic to get there: 18 GB/s over the first 24 uploads of a cold link, 53 GB/s after 32; profiles/README.md)
    e2e_run(max(32, (args.e2e_slots or 2 * LANES) + 1))
    barrier()
    t0e = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(dt.item())

    # ---- per-kernel durations (CUDA events on the launching stream), same workload ---------------
    kt = ev.kernel_times(dev_sets[0], d_lap, reps=max(3, min(args.steps, 10)))

    D = Dist(torch, dist, world, rank, dev)
    peak, peak_src = hbm_peak()
    # ---- merged top-10 against a host sort of every rank's lap times (NCCL path when N > 1) ----------------
    topk_same = topk_identity(D, ev, dev_sets[0], base, finish)
    # ---- parity against the reference-equivalent port (N = 1), both spline modes ---------------------------
    parity = None
    if cpu_result is not None:
        _rate, cpu_laps, cpu_a = cpu_result
        parity = {"sample": f"first {len(cpu_a)} rows of the timed population (seed 1002) against oracle/reference_port.py "
                            "(same SciPy FITPACK / numpy calls as the reference) on this host",
                  "numpy_pow": "SVML (AVX512 dispatch)" if "AVX512_SKX" in _numpy_simd() else "libm"}
        for mode in ("tridiagonal", "fitpack"):
            ev.set_spline_mode(mode)
            g = ev.lap_times(cpu_a)
            st_ = rel_stats(g, cpu_laps)
            st_["top10_identical"] = bool(np.array_equal(stable_topk(g, TOPK)[0], stable_topk(cpu_laps, TOPK)[0]))
            if cpu_laps_base is not None:  # the same rows against the reference's arithmetic on non-AVX512 hosts
                st_["vs_numpy_baseline_dispatch"] = rel_stats(g, cpu_laps_base)
            parity[mode] = st_
        if cpu_laps_base is not None:
            parity["reference_against_itself"] = dict(
                rel_stats(cpu_laps, cpu_laps_base),
                what="the port under numpy's dispatch on this host against the port under numpy's baseline dispatch (libm pow)")
        ev.set_spline_mode(args.spline)
    # ---- variants: FITPACK spline mode, fp32 sweeps -------------------------------------------------------
    variants = None
    if not args.no_variants and args.sweep_bits == 64 and args.spline == "tridiagonal":
        variants = {}
        vsteps = max(10, min(args.steps, 60))
        ev.set_spline_mode("fitpack")
        variants["fitpack_spline"] = short_run(D, ev, dev_sets, d_laps, vsteps, LANES, base, finish)
        ktf = ev.kernel_times(dev_sets[0], d_lap, reps=3)
        variants["fitpack_spline"]["kernel_ms"] = {k: round(v, 4) for k, v in ktf.items()}
        variants["fitpack_spline"]["note"] = ("LTK_SPLINE_FITPACK: SciPy FITPACK's fpclos / splder operation order, spline "
                                              "coefficients and derivatives bit-equal to the reference's")
        ev.set_spline_mode("tridiagonal")
        ev.set_sweep_precision(32)
        variants["fp32_sweeps"] = short_run(D, ev, dev_sets, d_laps, vsteps, LANES, base, finish)
        kt32 = ev.kernel_times(dev_sets[0], d_lap, reps=3)
        n_ = ns - 1
        b32 = 8 * na + 4 * n_ + 16 * n_ + 8  # K1b writes the fp32 curvature only; the sweeps read it twice, park 4-byte velocities
        variants["fp32_sweeps"].update({"kernel_ms": {k: round(v, 4) for k, v in kt32.items()}, "bytes_per_candidate": b32,
                                        "hbm_frac": b32 * B / (variants["fp32_sweeps"]["ms_per_step"] * 1e-3) / 1e9 / peak,
                                        "tolerance": "1e-4 relative to the fp64 kernels"})
        l32 = ev.lap_times_device(dev_sets[0]).cpu().numpy()
        ev.set_sweep_precision(64)
        l64 = ev.lap_times_device(dev_sets[0]).cpu().numpy()
        variants["fp32_sweeps"]["rel_err_vs_fp64"] = {k: v for k, v in rel_stats(l32, l64).items() if k != "count_over_1e-9"}
    # ---- BASELINE configs 3, 4, 5 at their named total sizes (strong-sharded) ----------------------------
    configs = None
    if not args.no_configs:
        configs = [
            run_config(D, ltk, local, "config 3: Buckmore + MX5, 1,048,576 candidates, top-10 all-gather", "MX5", None,
                       1 << 20, (2026, 3), False, peak, 3),
            run_config(D, ltk, local, "config 4: Bayesian-method database, 2^20 sampled trajectories (alphas, laps)", "tbr18",
                       None, 1 << 20, (2026, 4), True, peak, 3),
            run_config(D, ltk, local, "config 5: Buckmore resampled at 10,000 points per lap, 4,194,304 candidates", "tbr18",
                       10001, 1 << 22, (2026, 5), False, peak, 1),
            # the middle point of SURVEY 8(d)'s sampling-density sweep Ns in {846, 2,500, 10,000} (the ends are the
            # headline and config 5)
            run_config(D, ltk, local, "sampling-density sweep: 2,500 points per lap, 1,048,576 candidates", "tbr18",
                       2501, 1 << 20, (2026, 6), False, peak, 1),
        ]
    probe = h2d_probe(D, B * na * 8)

    if rank == 0:
        n = ns - 1
        # algorithmic bytes per candidate of each kernel: the shares of A_staged;
        # the K1a -> K1b hand-off (second derivatives, knots) is not in the model and not counted
        # fp32 variant: K1b evaluates and writes a 4-byte curvature, the sweeps read it and park 4-byte velocities
        sweep_bytes = 32 * n + 8 if args.sweep_bits == 64 else 16 * n + 8
        k1b_bytes = 8 * n if args.sweep_bits == 64 else 4 * n
        alg = {"k1a_spline_solve": 8 * na, "k1b_curvature": k1b_bytes, "k23_sweep": sweep_bytes}
        dom = max(kt, key=lambda k: kt[k])
        achieved = alg[dom] * B / (kt[dom] * 1e-3) / 1e9
        a_staged = 8 * na + k1b_bytes + sweep_bytes  # 8 Na + 40 n + 8 for fp64
        traffic, traffic_commit = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj_ = json.load(open(tp))
                traffic, traffic_commit = tj_.get(dom), tj_.get("_commit")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.sweep_bits == 64 else "f32 sweeps on f64 spline/curvature", "data": "synthetic",
            "config": dict(workload_config(args, na, ns),
                           kernels_per_step="k1a_solve, k1b_samples, k23_sweep (lap sum + top-10 in its epilogue)" +
                                            (" + topk_merge_gathered after the all-gather" if world > 1 else ""),
                           collective=("one all_gather of each rank's packed top-10 list (160 B) + one merge launch, on a "
                                       "communication stream" if world > 1 else "none")),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * na * 8,
                    "d2h_bytes_per_step": B * 8 + TOPK * 16,
                    "pipeline": f"LapTimeEvaluator.stream_populations: {LANES} populations in flight, each on its own "
                                f"compute stream; uploads and downloads on two copy streams; {args.e2e_slots or 2 * LANES} buffer sets"},
            "gpu_launches": launches,
            # `bound`: the contract's roof for this byte-moving path is HBM and `frac` is measured against it; what
            # actually limits the kernel is named in `binding` (FP64 issue + dependent-chain latency)
            "roofline": {"bound": "hbm", "binding": "fp64_issue", "kernel": dom, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_commit": traffic_commit,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_candidate": alg[dom],
                         "kernel_ms": {k: round(v, 4) for k, v in kt.items()},
                         # the roof that actually binds (DESIGN.md section 3 "What binds"): FP64-pipe cycles per
                         # sample from the SASS instruction counts and the measured issue costs (tools/ubench:
                         # 2 cycles per FP64 warp instruction and scheduler, 3 for a three-register DFMA)
                         "fp64_pipe": fp64_pipe_model(args, B, n, ms_total / args.steps, clocks),
                         "issue_model": issue_model(args, B, n, ms_total / args.steps, clocks),
                         # SURVEY 8(d) figure 2, the compulsory-traffic floor of a fully fused pipeline (alphas in,
                         # lap time out): how far from memory-bound the arithmetic is
                         "compulsory": {"bytes_per_candidate": 8 * na + 8,
                                        "frac": (8 * na + 8) * B / (ms_total / args.steps * 1e-3) / 1e9 / peak},
                         "pipeline": {"bytes_per_candidate": a_staged,
                                      "achieved": a_staged * B * world / (ms_total / args.steps * 1e-3) / 1e9 / world,
                                      "frac": a_staged * B / (ms_total / args.steps * 1e-3) / 1e9 / peak}},
        }
        line["topk_identical"] = topk_same
        line["h2d_probe"] = probe
        if cpu_result is not None:
            cores = os.cpu_count() or 1
            rate, cpu_laps, cpu_a = cpu_result
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "one_core": cpu_one_core,
                                    "sample": f"{args.cpu_sample} candidates of the timed population, warm fork pool of {cores} "
                                              "processes, oracle/reference_port.py",
                                    "parity_rel_err": parity[args.spline]}
            line["parity"] = parity
        if variants is not None:
            line["variants"] = variants
        if configs is not None:
            line["configs"] = configs
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
