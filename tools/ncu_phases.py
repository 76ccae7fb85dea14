#!/usr/bin/env python
"""Per-phase instruction table of a kernel from an .ncu-rep (SASS view): the SASS stream is cut at every BAR.SYNC
(and at extra marker opcodes given on the command line) and each segment's executed warp instructions are reported
per input base, with the opcode mix and the stall-sample share.  The segments sum to the kernel total.
usage: tools/ncu_phases.py REPORT.ncu-rep BASES_PER_LAUNCH [--dump]"""
import collections
import csv
import io
import subprocess
import sys

rep, bases = sys.argv[1], float(sys.argv[2])
dump = "--dump" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
I, S, SRC = ix["Instructions Executed"], ix["# Samples"], ix["Source"]
stall_cols = [(h, i) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h]
ALU = {"LOP3", "SHF", "PRMT", "ISETP", "VIMNMX", "SEL", "IADD3", "VIADD", "LEA", "R2P", "PLOP3", "VIADDMNMX", "MOV", "CS2R", "S2R", "P2R"}
LSU = {"LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "LDC", "LDCU", "ATOM", "SYNCS", "UBLKCP", "CCTL"}
CBU = {"BRA", "BSSY", "BSYNC", "BAR", "EXIT", "WARPSYNC", "NOP", "BREAK", "YIELD"}


def pipe(op):                       # issue class of an opcode (ALU and FMA each take a warp instruction every other cycle)
    if op.startswith("IMAD") or op in ("FFMA", "FMUL", "FADD"):
        return "fma"
    if op in ALU:
        return "alu"
    if op in LSU:
        return "lsu"
    if op in CBU:
        return "branch"
    return "other"


def new_seg(first):
    return {"n": 0, "s": 0, "ops": collections.Counter(), "first": first, "stalls": collections.Counter(), "lines": [],
            "pipe": collections.Counter()}


segs, cur = [], new_seg(0)
tot = ts = 0
for k, r in enumerate(rows[2:]):
    if len(r) <= I:
        continue
    try:
        e, s = int(r[I]), int(r[S])
    except ValueError:
        continue
    toks = r[SRC].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    cur["n"] += e; cur["s"] += s; cur["ops"][op.split(".")[0]] += e; cur["pipe"][pipe(op.split(".")[0])] += e
    cur["lines"].append((k, e, s, r[SRC].strip()))
    for h, i in stall_cols:
        try:
            cur["stalls"][h] += int(r[i])
        except ValueError:
            pass
    tot += e; ts += s
    if op.startswith("BAR") or op.startswith("EXIT"):
        segs.append(cur)
        cur = new_seg(k + 1)
segs.append(cur)
print(f"total warp instructions {tot:.4g} = {tot * 32 / bases:.2f} thread-instr/base; stall samples {ts}")
allp = collections.Counter()
for g in segs + [cur]:
    allp.update(g["pipe"])
print("by issue class (thread-instr/base): " + ", ".join(f"{c} {allp[c] * 32 / bases:.2f}" for c in ("alu", "fma", "lsu", "branch", "other")))
for j, g in enumerate(segs):
    if g["n"] == 0:
        continue
    ops = ", ".join(f"{o} {c * 32 / bases:.2f}" for o, c in g["ops"].most_common(8))
    st = ", ".join(f"{h[6:]} {c / max(1, ts):.3f}" for h, c in g["stalls"].most_common(4))
    pp = " ".join(f"{c} {g['pipe'][c] * 32 / bases:.2f}" for c in ("alu", "fma", "lsu", "branch", "other"))
    print(f"seg {j:2d} sass#{g['first']:5d} {g['n'] * 32 / bases:6.2f}/base stall {g['s'] / max(1, ts):5.3f} | {pp} | {ops} | {st}")
    if dump:
        for k, e, s, src in g["lines"]:
            if e:
                print(f"      {k:5d} {e * 32 / bases:6.3f} {s:6d}  {src}")
