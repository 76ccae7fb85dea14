#!/usr/bin/env python
"""End-to-end rate of s2k_run_packed2 (2-bit packed host input) on the C2 workload at several slab sizes.
Usage (GPU box): python tools/bench_packed.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import seq2kminmers_b200 as S

n_reads, read_len = 500_000, 20_000
n = n_reads * read_len
ctx = S.Context(0)
d_b = torch.empty(n, dtype=torch.uint8, device="cuda:0")
ctx.synth_device(1, 0, n, d_b.data_ptr())
torch.cuda.synchronize()
hb = torch.empty(n, dtype=torch.uint8).pin_memory(); hb.copy_(d_b); del d_b
hso = np.arange(n_reads + 1, dtype=np.uint64) * read_len
hp = torch.from_numpy(ctx.pack2(hb.numpy(), 16)).pin_memory()
hp_np = hp.numpy()
for slab_mib in (64, 128, 256, 512, 1024):
    ctx.set_slab_bytes(slab_mib << 20)
    ctx.run(hp_np, hso, 31, 5, 0.01, S.HashMode.HpcSimd, S.HashVariant.NT1_32, copy=False, packed2=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.run(hp_np, hso, 31, 5, 0.01, S.HashMode.HpcSimd, S.HashVariant.NT1_32, copy=False, packed2=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"slab={slab_mib} MiB: {n / dt / 1e9:.1f} Gbp/s ({dt * 1e3:.1f} ms per 10 Gbp)", flush=True)
