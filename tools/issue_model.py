#!/usr/bin/env python3
"""Issue-cycle cost of the straight-line blocks of one kernel under the model of DESIGN.md section 3 ("What binds"):
one cycle per instruction, two per FP64 instruction (DADD / DMUL / DFMA / DSETP), three per DFMA whose three source
operands are three DISTINCT vector registers (tools/ubench/fp64_operands.cu: 3.01 cycles against 2.01).  Needs cuobjdump.

    python tools/issue_model.py lap_time_optimization_b200/libltk.so k23_sweepILi0ELi0E [min_block_size=40]

Prints, for every block of at least `min_block_size` instructions: instructions, FP64 ones, three-operand DFMAs, model
cycles.  The figures bench.py's `roofline.issue_model` uses come from here: K23's regular phase-1 / phase-2 blocks of two
row pairs plus the loop header in front of them, K1b's sample loop."""
import re
import subprocess
import sys

FP64 = ("DADD", "DMUL", "DFMA", "DSETP")
ENDS = ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BAR", "WARPSYNC", "BRX", "JMP")


def distinct_vector_sources(operands):
    """Number of distinct vector registers among the SOURCE operands of an FP64 instruction (destination first)."""
    ops = [o.strip() for o in operands.split(",")][1:]
    regs = set()
    for o in ops:
        m = re.search(r"\bR(\d+)\b", o)
        if m and not o.lstrip("-|~!").startswith(("UR", "c[")):
            regs.add(int(m.group(1)))
    return len(regs)


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    min_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    body = next(f for f in funcs if f.startswith("_Z") and pat in f.split("\n")[0])
    print("function", body.split("\n")[0])
    block, blocks = [], []
    for ln in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P[0-9T]+\s+)?)([A-Z0-9_.]+)\s*(.*?);", ln)
        if not m:
            continue
        addr, op, rest = int(m.group(1), 16), m.group(3), m.group(4)
        block.append((addr, op, rest))
        if op.split(".")[0] in ENDS:
            blocks.append(block)
            block = []
    print(f"{'address':>9s} {'instr':>6s} {'fp64':>5s} {'dfma3':>6s} {'cycles':>7s}  ends with")
    for b in blocks:
        if len(b) < min_n:
            continue
        n_fp64 = n3 = 0
        for _, op, rest in b:
            root = op.split(".")[0]
            if root in FP64:
                n_fp64 += 1
                if root == "DFMA" and distinct_vector_sources(rest) >= 3:
                    n3 += 1
        cycles = len(b) + n_fp64 + n3
        print(f"{b[0][0]:#9x} {len(b):6d} {n_fp64:5d} {n3:6d} {cycles:7d}  {b[-1][1]} {b[-1][2][:30]}")


if __name__ == "__main__":
    main()
