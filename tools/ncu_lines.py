#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, opcode mix and thread-instructions per base per source line.
usage: tools/ncu_lines.py REPORT.ncu-rep BASES_PER_LAUNCH [top]"""
import collections
import csv
import io
import subprocess
import sys

rep, bases = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30


def page(*args):
    return subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(page("--page", "raw"))))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:75s} {vals[i]:>16s} {units[i]}")

rows = list(csv.reader(io.StringIO(page("--page", "source", "--print-source", "cuda,sass"))))
h = rows[2]
ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
src = None
for r in rows[:2]:
    if r and r[0] == "File Path":
        try:
            src = open(r[1]).read().split("\n")
        except OSError:
            src = None
per, samp, ops, cur = collections.Counter(), collections.Counter(), collections.Counter(), None
for r in rows[3:]:
    if r and r[0].strip().isdigit():
        cur = int(r[0])
        continue
    if cur is None or len(r) <= ie:
        continue
    try:
        e, s = int(r[ie]), int(r[isamp])
    except ValueError:
        continue
    per[cur] += e
    samp[cur] += s
    toks = r[3].split()
    if toks:
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        ops[op.split(".")[0]] += e
tot, ts = sum(per.values()), max(1, sum(samp.values()))
print(f"\nwarp instructions {tot:.4g} -> {tot * 32 / bases:.1f} thread-instr/base")
print("opcodes: " + ", ".join(f"{o} {c / tot:.3f}" for o, c in ops.most_common(16)))
for ln, e in per.most_common(top):
    text = src[ln - 1].strip()[:100] if src and ln - 1 < len(src) else ""
    print(f"{e / tot:6.3f} {e * 32 / bases:6.2f}/base stall={samp[ln] / ts:5.3f} L{ln:4d} {text}")
