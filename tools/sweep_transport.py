#!/usr/bin/env python
"""Sweep the host->device transport knobs of s2k_run (pack ratio, packer threads, slab size) on the C2 workload.
Usage (GPU box): python tools/sweep_transport.py > gpurun_out/sweep_transport.txt"""
import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import seq2kminmers_b200 as S

def main():
    n_reads, read_len = 500_000, 20_000
    n = n_reads * read_len
    ctx = S.Context(0)
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    ctx.synth_device(1, 0, n, d_b.data_ptr())
    torch.cuda.synchronize()
    hb = torch.empty(n, dtype=torch.uint8).pin_memory(); hb.copy_(d_b); del d_b
    hso = np.arange(n_reads + 1, dtype=np.uint64) * read_len
    hb_np = hb.numpy()
    print("host cores", os.cpu_count(), flush=True)
    def run():
        return ctx.run(hb_np, hso, 31, 5, 0.01, S.HashMode.HpcSimd, S.HashVariant.NT1_32, copy=False)
    for slab_mib, thr, ratio in itertools.product((256,), (10, 12, 14), (0.0, 0.5, 0.6, 0.7, 0.8)):
        if ratio == 0.0 and thr != 10: continue
        ctx.set_slab_bytes(slab_mib << 20); ctx.set_transport(thr, ratio)
        run(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): out = run()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"slab={slab_mib}MiB threads={thr} ratio={ratio}: {n/dt/1e9:.1f} Gbp/s  transport={ctx.last_transport()}", flush=True)
main()
