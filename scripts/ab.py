#!/usr/bin/env python3
"""A/B of library builds on the GPU box: bench.py (no CPU baseline / configs / variants) once per build and
round, builds interleaved so that box-to-box and power-state drift hits all of them alike.

    python scripts/ab.py [--rounds 3] [--steps 300] [--args "--lanes 3"] default la0_9 la0_10 ...

`name@VAR=value,...` sets environment knobs for that run (scripts/README.md); `default` is lap_time_optimization_b200/libltk.so, any other name lap_time_optimization_b200/variants/libltk_<name>.so
(`make -C lap_time_optimization_b200/csrc variant NAME=<name> EXTRA='-D...'`, loaded through LTK_LIB_PATH)."""
import argparse
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--rounds", type=int, default=3)
    p.add_argument("--steps", type=int, default=300)
    p.add_argument("--args", default="")
    p.add_argument("names", nargs="+")
    a = p.parse_args()
    res = {n: [] for n in a.names}
    for r in range(a.rounds):
        for n in a.names:
            env = dict(os.environ)
            lib_name, _, extra = n.partition("@")  # name@VAR=value,VAR=value: environment knobs of the library
            for kv in filter(None, extra.split(",")):
                k, _, v = kv.partition("=")
                env[k] = v
            n_lib = lib_name
            if n_lib != "default":
                env["LTK_LIB_PATH"] = os.path.join(ROOT, "lap_time_optimization_b200", "variants", f"libltk_{n_lib}.so")
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(a.steps), "--warmup", "5",
                   "--no-cpu-baseline", "--no-configs", "--no-variants"] + a.args.split()
            out = subprocess.run(cmd, env=env, capture_output=True, text=True)
            line = next((ln for ln in out.stdout.splitlines() if ln.startswith("{")), None)
            if line is None:
                print(n, "FAILED", out.stderr[-400:], flush=True)
                continue
            d = json.loads(line)
            km = d["roofline"]["kernel_ms"]
            res[n].append((d["ms_per_step"], d["e2e"]["value"] / 1e6, km["k1a_spline_solve"], km["k1b_curvature"],
                           km["k23_sweep"], d["clocks"]["sm_mhz"]))
            print(f"{n:12s} r{r} step {d['ms_per_step']:.4f} ms  e2e {d['e2e']['value'] / 1e6:6.1f} M  "
                  f"k1a {km['k1a_spline_solve']:.4f} k1b {km['k1b_curvature']:.4f} k23 {km['k23_sweep']:.4f}  "
                  f"{d['clocks']['sm_mhz']:.0f} MHz", flush=True)
    print("--- medians (step ms, e2e M/s, k1a, k1b, k23, MHz)")
    for n in a.names:
        if res[n]:
            print(f"{n:12s}", "  ".join(f"{statistics.median(c):.4f}" for c in zip(*res[n])))


if __name__ == "__main__":
    main()
