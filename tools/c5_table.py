#!/usr/bin/env python
"""Formats the JSON lines of `bench.py --workload c5` as the table of profiles/*_config5_sweep.md.
usage: tools/c5_table.py sweep.jsonl > table.md"""
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{") and '"density"' in l]
ks = sorted({r["k"] for r in rows})
print("| HPC | density | " + " | ".join(f"k={k}" for k in ks) + " | CPU port (all host threads) | parity (digest of every read of slab 1 at k = 5 and 10) |")
print("|---|---|" + "---|" * (len(ks) + 2))
for hpc in (True, False):
    for d in sorted({r["density"] for r in rows}):
        sel = {r["k"]: r for r in rows if r["hpc"] == hpc and r["density"] == d}
        if not sel:
            continue
        par = [r["parity"] for r in sel.values() if r.get("parity")]
        ptxt = ("match (%d reads x %d corners)" % (par[0]["reads_checked"], len(par))) if par and all(p["digest_match"] for p in par) else ("MISMATCH" if par else "-")
        cpu = next(iter(sel.values())).get("cpu_port_Gbp_per_s")
        print(f"| {'on' if hpc else 'off'} | {d} | " + " | ".join(f"{sel[k]['Gbp_per_s']:.0f}" if k in sel else "" for k in ks) + f" | {cpu:.1f} | {ptxt} |")
