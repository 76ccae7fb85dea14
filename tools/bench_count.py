#!/usr/bin/env python
"""Consumer-side counting on config-2 shaped reads: items/s of the device table, and -- under torchrun
-- of the hash-partitioned multi-GPU form with its NCCL all-to-all.  Usage: tools/bench_count.py [reads_per_gpu]"""
import importlib
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
S = importlib.import_module("rust-seq2kminmers_b200")
C = importlib.import_module("rust-seq2kminmers_b200.counting")
sharding = importlib.import_module("rust-seq2kminmers_b200.sharding")

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
L, seed = 20000, 0x5EED0002
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = S.Context(local)
n_bases = n_reads * L
d_bases = torch.empty(n_bases + 16, dtype=torch.uint8, device=dev)
# every rank takes the SAME stream slice shifted by half a batch: half of each rank's reads recur on its neighbour
ctx.synth_device(seed, (rank * n_reads // 2) * L, n_bases, d_bases.data_ptr())
d_so = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * L
stream = torch.cuda.current_stream().cuda_stream
res = ctx.run_device(d_bases.data_ptr(), d_so.data_ptr(), n_reads, n_bases, 31, 5, 0.01, S.HashMode.HpcSimd, stream=stream,
                     no_minimizer_stream=True)
n_items = int(res.n_items)
first = 0
if world > 1:
    _, first = sharding.gather_totals(n_items, int(res.n_minimizers), device=dev)
out = {"n_gpus": world, "items_per_gpu": n_items}
for rep in range(3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if world == 1:
        r = ctx.count_device(res.hash, n_items, 0, 0, stream)
        nd, t_x = int(r.n_distinct), 0.0
    else:
        r, _, _, t_x = C.count_distributed(ctx, res.hash, n_items, first, dev, fetch=False)
        nd = int(r.n_distinct)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
t = torch.tensor([dt, t_x], dtype=torch.float64, device=dev)
s = torch.tensor([nd, n_items], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s)
if rank == 0:
    out.update(seconds=float(t[0]), exchange_seconds=float(t[1]), distinct=int(s[0]), items=int(s[1]),
               items_per_s=int(s[1]) / float(t[0]),
               note="count_device = insert + compact kernels (last of 3 repeats); multi-GPU adds bucket-by-hash, the NCCL "
                    "all-to-all of (hash, id) pairs; the table stays on the device")
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
ctx.close()
