#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --set full` report: DRAM bytes (read + write) per launch of each pipeline
kernel, averaged over the captured launches, stamped with the commit the capture was taken at.  bench.py copies the
dominant kernel's figure into `roofline.traffic` and the stamp into `roofline.traffic_commit`.

    python tools/ncu_traffic.py gpurun_out/r2k_full.ncu-rep profiles/traffic.json [commit]"""
import csv
import json
import subprocess
import sys

NAMES = {"k1a_solve": "k1a_spline_solve", "k1a_fitpack": "k1a_fitpack", "k1b_samples": "k1b_curvature",
         "k23_sweep": "k23_sweep", "k23_roles": "k23_roles", "k23_f32": "k23_f32", "topk_select": "topk_select"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    commit = sys.argv[3] if len(sys.argv) > 3 else subprocess.run(
        ["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik, ir, iw, it = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                             "gpu__time_duration.sum"))
    acc = {}
    for r in data:
        key = next((v for k, v in NAMES.items() if k in r[ik]), None)
        if key is None:
            continue
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        acc.setdefault(key, []).append((b, float(r[it])))
    res = {"_commit": commit,
           "_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches) from one "
                    f"`ncu --set full` capture ({rep.split('/')[-1]}) of `bench.py` at 65,536 candidates, one population "
                    "at a time; `_us` = gpu__time_duration of the same launches under ncu (cold, serialised)"}
    for k, v in acc.items():
        res[k] = sum(x[0] for x in v) / len(v)
        res[k + "_us"] = round(sum(x[1] for x in v) / len(v), 2)
        res[k + "_launches"] = len(v)
    if "k23_sweep" in res and "k23_roles" in res:
        # a single population's sweep runs split (whole layers on k23_sweep, the remainder on k23_roles, concurrently):
        # bench.py times the whole population's sweep, so `k23_sweep` is the sum of the two parts
        res["k23_sweep_split_main"] = res["k23_sweep"]
        res["k23_sweep_split_remainder_k23_roles"] = res.pop("k23_roles")
        res["k23_sweep"] = res["k23_sweep_split_main"] + res["k23_sweep_split_remainder_k23_roles"]
        res["_note"] += "; k23_sweep = the whole population's sweep = k23_sweep (56,832 candidates) + k23_roles (8,704)"
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
