#!/usr/bin/env python
"""Resident throughput of the H = u64 and H = u16 flavours (S2K_HASH_NT1_64 / S2K_HASH_NT1_16, modes Hpc and Regular) on a config-2 shaped batch
(reads x 20 kb, l=31 k=5 d=0.01), next to the 32-bit scalar-profile run of the same mode.  CUDA events on the stream.
Usage (GPU box): python tools/bench_h64.py [reads]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import seq2kminmers_b200 as S

n_reads, read_len = (int(sys.argv[1]) if len(sys.argv) > 1 else 250_000), 20_000
n = n_reads * read_len
ctx = S.Context(0)
d_b = torch.empty(n + 64, dtype=torch.uint8, device="cuda:0")
ctx.synth_device(0x5EED0002, 0, n, d_b.data_ptr())
d_so = torch.arange(n_reads + 1, dtype=torch.int64, device="cuda:0") * read_len
torch.cuda.synchronize()
for mode in (S.HashMode.Hpc, S.HashMode.Regular):
    for var in (S.HashVariant.NT1_32, S.HashVariant.NT1_64, S.HashVariant.NT1_16):
        for _ in range(2):
            r = ctx.run_device(d_b.data_ptr(), d_so.data_ptr(), n_reads, n, 31, 5, 0.01, mode, var)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = ctx.run_device(d_b.data_ptr(), d_so.data_ptr(), n_reads, n, 31, 5, 0.01, mode, var)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{mode.name:8s} {var.name}: {n / ms / 1e6:7.1f} Gbp/s  ({ms:.2f} ms per {n / 1e9:.1f} Gbp, {r.n_items} items, {r.n_minimizers} minimizers)", flush=True)
