"""Developer probe (GPU box): parity statistics and kernel timings, printed, nothing asserted."""
import glob, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lap_time_optimization_b200 as ltk
from oracle.c_oracle import COracle
from oracle.reference_port import OracleTrack, load_vehicle, top_k

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    print(torch.cuda.get_device_name(0))
    for tag in sorted(glob.glob(os.path.join(ROOT, "tests/golden/*.npz"))):
        g = np.load(tag); name = os.path.basename(tag)[:-4]
        tname = name.split("_")[0]
        width = 1.0 if "w100" in name else (0.6 if ("full" in name and tname != "buckmore") else 0.8)
        vj = ltk.data_path("vehicles", "MX5.json" if "mx5" in name else "tbr18.json")
        tj = ltk.data_path("tracks", tname + ".json")
        mode = "full" if "full" in name else "bayes"
        ns = int(g["ns"])
        ev = ltk.LapTimeEvaluator(ltk.Track(tj, track_width=width, quiet=True), ltk.load_vehicle(vj), mode, ns)
        laps = ev.lap_times(g["alphas"])
        co = COracle(OracleTrack(tj, width), load_vehicle(vj), mode, ns, device_sum_order=True)
        cl = co.lap_times(g["alphas"])
        rel = np.abs(laps - g["laps"]) / g["laps"]
        pr = ev.profile(g["alphas"][0]); cp = co.profile(g["alphas"][0])
        eq = {k: bool(np.array_equal(pr[k], cp[k])) for k in ("k", "v_local", "v_acclim", "v_declim", "v")}
        print(f"{name:32s} vs C-oracle biteq {np.sum(laps == cl)}/{len(laps)} maxrel {np.max(np.abs(laps-cl)/cl):.1e} | vs golden med {np.median(rel):.1e} max {rel.max():.1e} | profile biteq {eq} lap {pr['lap']==cp['lap']}")
        ev.close()
    # big batch
    tj, vj = ltk.data_path("tracks", "buckmore.json"), ltk.data_path("vehicles", "tbr18.json")
    for veh in ("tbr18.json", "MX5.json"):
        vj = ltk.data_path("vehicles", veh)
        ev = ltk.LapTimeEvaluator(ltk.Track(tj, track_width=0.8, quiet=True), ltk.load_vehicle(vj), "bayes")
        a = np.random.default_rng(1002).uniform(0, 0.99, (B, ev.n_alpha))
        d_a = torch.as_tensor(a).cuda()
        d_lap = ev.lap_times_device(d_a); torch.cuda.synchronize()
        for g_over in (None,):
            ts = []
            for it in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); d_lap = ev.lap_times_device(d_a); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(veh, "B", B, "ms", [f"{t:.3f}" for t in ts], "evals/s %.3e" % (B / (min(ts) * 1e-3)))
        laps = d_lap.cpu().numpy()
        t = time.time(); cl = COracle(OracleTrack(tj, 0.8), load_vehicle(vj), "bayes", device_sum_order=True).lap_times(a); tc = time.time() - t
        print("  C oracle %.1f evals/s; biteq %d/%d maxrel %.2e" % (B / tc, np.sum(laps == cl), B, np.max(np.abs(laps - cl) / cl)))
        best, idx = ev.topk(d_lap, 10)
        oi, ob = top_k(list(cl), 10)
        print("  topk equal", np.array_equal(idx, oi), np.array_equal(best, ob), idx[:4])
        ev.close()

main()
