#!/usr/bin/env python3
"""Warp-stall hot spots of one kernel from an `ncu --set full --import-source on` report (run where ncu is).

    python tools/ncu_stalls.py gpurun_out/r1t_full.ncu-rep k23_sweep [N]

Prints the stall reasons' shares of all sampled warp states and the N SASS instructions that collected the
most samples, each with its two leading reasons -- the view that showed K23's look-ahead loads (long
scoreboard on the moves that consume them) and the two serial load latencies at the start of every K1b CTA."""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat,
                          "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen, data = set(), []
    for r in rows[2:]:  # the export repeats every instruction; keep the first copy
        if len(r) == len(hdr) and r[ix["# Samples"]].isdigit() and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    total = sum(int(r[ix["# Samples"]]) for r in data)
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(rows[0][1] if len(rows[0]) > 1 else pat)
    print("samples", total, "instructions", len(data))
    for k in sorted(reasons, key=lambda k: -sum(int(r[ix[k]]) for r in data))[:8]:
        v = sum(int(r[ix[k]]) for r in data)
        print(f"  {k:24s}{v:8d} {100 * v / total:5.1f} %")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
        lead = sorted(((k, int(r[ix[k]])) for k in reasons), key=lambda x: -x[1])[:2]
        n = int(r[ix["# Samples"]])
        print(f"{n:6d} {100 * n / total:4.1f} % x{r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:58]:58s} {lead}")


if __name__ == "__main__":
    main()
