#!/usr/bin/env python3
"""Basic blocks of one kernel in a cubin/.so in address order (needs cuobjdump): size, FP64 / load / store /
vote counts and the instruction that ends each block -- enough to add up what one loop iteration executes.

    python tools/sass_blocks.py lap_time_optimization_b200/libltk.so k23_sweepILi0ELi0 [0x1600 0x5600]"""
import collections
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    body = next(f for f in funcs if f.startswith("_Z") and pat in f.split("\n")[0])
    ops = []
    for ln in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P[0-9T]+\s+)?)([A-Z0-9_.]+)(.*?);", ln)
        if m:
            ops.append((int(m.group(1), 16), m.group(3), m.group(2).strip(), m.group(4).strip()))
    blocks, cur = [], []
    for o in ops:
        cur.append(o)
        if o[1].split(".")[0] in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BAR", "WARPSYNC", "BRX", "JMP"):
            blocks.append(cur)
            cur = []
    for b in blocks:
        if not lo <= b[0][0] <= hi:
            continue
        h = collections.Counter(o[1].split(".")[0] for o in b)
        fp64 = sum(h[k] for k in ("DADD", "DMUL", "DFMA", "DSETP"))
        last = b[-1]
        print(f"{b[0][0]:#07x} n={len(b):4d} fp64={fp64:3d} ldg={h['LDG']} stg={h['STG']} vote={h['VOTE']} "
              f"ends {last[2]} {last[1]} {last[3][:40]}")


if __name__ == "__main__":
    main()
