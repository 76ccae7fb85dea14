#!/usr/bin/env python
"""Per-phase cycle breakdown of k_minimizers (thread 0 of every CTA, summed).  Needs a library built with
S2K_NVCC_EXTRA=-DS2K_PHASE_CLOCKS.  Usage (GPU box): python tools/phase_clocks.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import seq2kminmers_b200 as S

NAMES = ["ticket+clear", "load+starts (S2,S3a)", "keep mask+scan (S3b)", "compaction (S4)", "flags+walk (S4b)",
         "hash+hit scan (S5,S6a)", "claim (S6b)", "emission (S7)", "seq offsets (S8)"]
ctx = S.Context(0)
n_reads, read_len = 50_000, 20_000
n = n_reads * read_len
d_b = torch.empty(n, dtype=torch.uint8, device="cuda:0")
ctx.synth_device(1, 0, n, d_b.data_ptr())
d_so = torch.arange(n_reads + 1, dtype=torch.int64, device="cuda:0") * read_len
torch.cuda.synchronize()
fn = ctx.lib.c.s2k_debug_phase_clocks
out = (C.c_ulonglong * 16)()
for rep in range(2):
    fn(out)
    ctx.run_device(d_b.data_ptr(), d_so.data_ptr(), n_reads, n, 31, 5, 0.01, S.HashMode.HpcSimd, S.HashVariant.NT1_32)
    torch.cuda.synchronize()
fn(out)
tot = sum(out[i] for i in range(9))
for i, nm in enumerate(NAMES):
    print(f"{nm:28s} {out[i] / tot * 100:6.2f} %   {out[i] / (n / 16128):10.0f} cycles/tile")
print("total cycles/tile", tot / (n / 16128))
