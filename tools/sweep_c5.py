#!/usr/bin/env python
"""BASELINE config 5: synthetic HiFi reads (20 kb), sweep density {0.001,0.002,0.005,0.01} x k {5..10} x HPC {on,off},
device-resident, one GPU (run under torchrun for more: every rank sweeps its own reads; rank 0 prints the aggregate).
Usage: tools/sweep_c5.py [reads_per_gpu] -> one JSON line per point + a markdown table on stderr."""
import importlib
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
S = importlib.import_module("rust-seq2kminmers_b200")

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000          # 10 Gbp per GPU (the config's 30 Gbp needs 3 slabs)
L, seed = 20000, 0x5EED0005
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = S.Context(local)
n_bases = n_reads * L
d_bases = torch.empty(n_bases + 16, dtype=torch.uint8, device=dev)
ctx.synth_device(seed, rank * n_bases, n_bases, d_bases.data_ptr())
d_so = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * L
stream = torch.cuda.current_stream().cuda_stream
rows = []
for hpc in (True, False):
    mode = S.HashMode.HpcSimd if hpc else S.HashMode.Simd
    for density in (0.001, 0.002, 0.005, 0.01):
        for k in range(5, 11):
            run = lambda: ctx.run_device(d_bases.data_ptr(), d_so.data_ptr(), n_reads, n_bases, 31, k, density, mode, stream=stream)
            for _ in range(2):
                res = run()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                res = run()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
            c = torch.tensor([int(res.n_items), int(res.n_minimizers)], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(c, op=dist.ReduceOp.SUM)
            ms = float(t.item())
            items = int(c[0].item())
            row = {"hpc": hpc, "density": density, "k": k, "n_gpus": world, "Gbp": world * n_bases / 1e9, "ms": ms,
                   "Gbp_per_s": world * n_bases / ms / 1e6, "items": items, "minimizers": int(c[1].item()),
                   "bytes_per_base": (world * n_bases + 8 * world * (n_reads + 1) + 17 * items) / (world * n_bases)}
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
if rank == 0:
    print("| HPC | density | " + " | ".join(f"k={k}" for k in range(5, 11)) + " |", file=sys.stderr)
    print("|---|---|" + "---|" * 6, file=sys.stderr)
    for hpc in (True, False):
        for density in (0.001, 0.002, 0.005, 0.01):
            vals = [r["Gbp_per_s"] for r in rows if r["hpc"] == hpc and r["density"] == density]
            print(f"| {'on' if hpc else 'off'} | {density} | " + " | ".join(f"{v:.0f}" for v in vals) + " |", file=sys.stderr)
if world > 1:
    dist.destroy_process_group()
