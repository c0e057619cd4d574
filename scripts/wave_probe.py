"""Developer probe (GPU box): throughput of a large population scored in chunks of different sizes over
`lanes` lanes (LapTimeEvaluator._lap_times_waves), for a given sampling density.
    python scripts/wave_probe.py NS TOTAL CHUNK[,CHUNK...] [LANES[,LANES...]]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import lap_time_optimization_b200 as ltk  # noqa: E402


def main():
    ns, total = int(sys.argv[1]), int(sys.argv[2])
    chunks = [int(x) for x in sys.argv[3].split(",")]
    lanes = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "3").split(",")]
    tj, vj = ltk.data_path("tracks", "buckmore.json"), ltk.data_path("vehicles", "tbr18.json")
    for nl in lanes:
        for chunk in chunks:
            ev = ltk.LapTimeEvaluator(ltk.Track(tj, track_width=0.8, quiet=True), ltk.load_vehicle(vj), "bayes", ns, device=0)
            ev.WAVE, ev.wave_lanes = chunk, nl
            d_a = ev.random_population_device(total, (1, 2))
            out = torch.empty(total, dtype=torch.float64, device="cuda")
            ev.lap_times_device(d_a[:min(total, nl * chunk)], out=out[:min(total, nl * chunk)])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ev.lap_times_device(d_a, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            n = ns - 1
            print(f"ns {ns} total {total} chunk {chunk} lanes {nl}: {ms:.2f} ms, {total / ms * 1e-3:.3f} M evals/s, "
                  f"HBM {(8 * 43 + 40 * n + 8) * total / ms / 1e6 / 6545.6:.3f}", flush=True)
            ev.close()
            del d_a, out
            torch.cuda.empty_cache()


main()
