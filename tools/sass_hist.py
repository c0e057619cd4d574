#!/usr/bin/env python3
"""Opcode histogram of the largest straight-line blocks of one kernel in a cubin/.so (needs cuobjdump).

    python tools/sass_hist.py lap_time_optimization_b200/libltk.so k23_sweepILi0ELi0 [nblocks]

Splits the SASS at branches / barriers / calls and prints, for the N longest blocks, the instruction
count by opcode and by pipe class (FP64 = DADD/DMUL/DFMA/DSETP; XU = MUFU/I2F/F2I/F2F)."""
import collections
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    nb = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    body = next(f for f in funcs if f.startswith("_Z") and pat in f.split("\n")[0])
    print("function", body.split("\n")[0])
    ops = []
    for ln in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            ops.append((int(m.group(1), 16), m.group(2)))
    blocks, cur = [], []
    for addr, op in ops:
        cur.append((addr, op))
        root = op.split(".")[0]
        if root in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BAR", "WARPSYNC", "BRX", "JMP"):
            blocks.append(cur)
            cur = []
    if cur:
        blocks.append(cur)
    print("instructions", len(ops), "blocks", len(blocks))
    for blk in sorted(blocks, key=len, reverse=True)[:nb]:
        h = collections.Counter(op.split(".")[0] for _, op in blk)
        fp64 = sum(h[k] for k in ("DADD", "DMUL", "DFMA", "DSETP"))
        xu = sum(h[k] for k in ("MUFU", "I2F", "F2I", "F2F"))
        print(f"-- block @{blk[0][0]:#x}: {len(blk)} instr, FP64 {fp64}, XU {xu}, other {len(blk) - fp64 - xu}")
        print("   " + ", ".join(f"{k} {v}" for k, v in h.most_common()))


if __name__ == "__main__":
    main()
