#!/usr/bin/env python3
"""In-order issue model of one straight-line SASS block (lone warp): when does each instruction issue
given B200 latencies measured by tools/ubench (FP64 8.4 cycles dependent, MUFU 17, I2F 23, LDS ~30),
and how long does the block take?  Shows how well the compiler interleaved independent chains.

    python tools/sass_sim.py lib.so <function substring> <hex start addr> [n instr]
"""
import re
import subprocess
import sys

LAT = {"DFMA": 8.4, "DMUL": 8.4, "DADD": 8.4, "DSETP": 12, "MUFU": 17.5, "I2F": 23, "LDS": 30, "LDG": 600, "LDCU": 0,
       "FSEL": 5, "SEL": 5, "ISETP": 6, "IMAD": 5, "VIADD": 5, "IADD3": 5, "LEA": 5, "SHF": 5, "VIMNMX": 5,
       "VIADDMNMX": 5, "MOV": 5, "LOP3": 5, "PLOP3": 6, "S2UR": 0, "UMOV": 0, "ULEA": 0}
FP64 = {"DFMA", "DMUL", "DADD", "DSETP"}


def regs(tok):
    out = []
    for m in re.finditer(r"\b(R\d+|P\d+|UR\d+|UP\d+)(?:\.64)?", tok):
        out.append(m.group(1))
    return out


def main():
    lib, pat, start = sys.argv[1], sys.argv[2], int(sys.argv[3], 16)
    count = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    body = next(f for f in funcs if pat in f.split("\n")[0])
    ins = []
    for ln in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);", ln)
        if not m:
            continue
        addr = int(m.group(1), 16)
        if addr < start:
            continue
        op = m.group(3).split(".")[0]
        if op in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BAR", "WARPSYNC") or len(ins) >= count:
            break
        ops = m.group(4).split(",")
        wide = op in FP64 or op in ("MUFU", "I2F") or ".64" in m.group(3) or ".128" in m.group(3)
        dst = regs(ops[0]) if ops else []
        src = [r for o in ops[1:] for r in regs(o)] + (regs(m.group(2)) if m.group(2) else [])
        if op in ("STG", "STS", "ST"):
            src, dst = dst + src, []
        def widen(rs):
            res = []
            for r in rs:
                res.append(r)
                if r[0] == "R" and wide:
                    res.append("R%d" % (int(r[1:]) + 1))
            return res
        ins.append((addr, op, widen(dst), widen(src), m.group(3)))
    ready = {}
    t = 0.0
    fp64_busy = 0.0
    nfp = 0
    for addr, op, dst, src, full in ins:
        issue = t
        for r in src:
            issue = max(issue, ready.get(r, 0.0))
        if op in FP64:
            issue = max(issue, fp64_busy)
            vec = {r for r in src if r[0] == "R" and int(r[1:]) % 2 == 0}
            cost = 3.0 if (op == "DFMA" and len(vec) >= 3) else 2.0
            fp64_busy = issue + cost
            nfp += 1
        lat = LAT.get(op, 5)
        for r in dst:
            ready[r] = issue + lat
        t = issue + 1.0
    print(f"{len(ins)} instructions, {nfp} FP64; lone-warp in-order time {t:.0f} cycles "
          f"({t / max(len(ins), 1):.2f} cycles/instr)")


if __name__ == "__main__":
    main()
