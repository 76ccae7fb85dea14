#!/usr/bin/env python
"""Throughput of the standalone RLE entry point (s2k_encode_rle -> k_rle; `encode_rle_simd` of src/hpc.rs:44-147 over a
batch): kernel time by CUDA events (s2k_ctx_set_timing) and the whole call with host buffers in and out, checked against
numpy.  No torch.  Usage (GPU box): python tools/bench_rle.py [reads] [read_len]   (S2K_LIB=path: another library)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import seq2kminmers_b200 as S
from oracle import oracle as O

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
n = n_reads * L
bases = O.synth(0x5EED0002, 0, n)
so = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(L))
keep = np.ones(n, dtype=bool)
keep[1:] = bases[1:] != bases[:-1]
keep[::L] = True                                           # the first base of every read is kept
lib = S.Library(os.environ["S2K_LIB"]) if os.environ.get("S2K_LIB") else None
ctx = S.Context(0, lib) if lib else S.Context(0)
ctx.set_timing(True)
import ctypes as C
def pinned_copy(a):                                        # the caller's buffers in pinned memory (s2k_host_alloc)
    ptr = C.c_void_p()
    assert ctx.lib.c.s2k_host_alloc(a.nbytes, C.byref(ptr)) == 0
    v = np.frombuffer((C.c_uint8 * a.nbytes).from_address(ptr.value), dtype=a.dtype)
    v[:] = a
    return v
pb, pso = pinned_copy(bases), pinned_copy(so)
for scalar_rule in (False, True):
    best_k, best_w = 1e30, 1e30
    for rep in range(3):
        t0 = time.perf_counter()
        h, p, off = ctx.encode_rle(bases, so, scalar_rule=scalar_rule)
        w = time.perf_counter() - t0
        k = ctx.last_kernel_ms()[0]
        best_k, best_w = min(best_k, max(k, 1e-9)), min(best_w, w)
    best_c = 1e30
    res = S._RleResult()
    ctx.lib.c.s2k_ctx_set_flags(ctx.h, 4 if scalar_rule else 0)
    for rep in range(3):                                   # the C ABI call alone: pinned buffers in, pinned results out
        t0 = time.perf_counter()
        assert ctx.lib.c.s2k_encode_rle(ctx.h, pb.ctypes.data, pso.ctypes.data, n_reads, C.byref(res)) == 0
        best_c = min(best_c, time.perf_counter() - t0)
    ctx.lib.c.s2k_ctx_set_flags(ctx.h, 0)
    assert res.n_hpc == int(keep.sum())
    assert np.array_equal(h, bases[keep]) and np.array_equal(p, (np.flatnonzero(keep) % L).astype(np.uint32))
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(keep.reshape(n_reads, L).sum(1))]).astype(np.uint64))
    # bytes the kernel must move: 1 read per base, 1 + 4 written per kept base, offsets
    alg = n + 5 * int(keep.sum()) + 16 * (n_reads + 1)
    print(f"k_rle {'scalar rule' if scalar_rule else 'simd rule  '}: {n / 1e9:.2f} Gbp, {int(keep.sum())} kept; kernel {best_k:.3f} ms = "
          f"{n / best_k / 1e6:.0f} Gbp/s, {alg / best_k / 1e6:.0f} GB/s algorithmic; s2k_encode_rle with pinned buffers "
          f"({(n + 5 * int(keep.sum())) / 1e9:.2f} GB over PCIe) {best_c * 1e3:.0f} ms = {n / best_c / 1e9:.1f} Gbp/s; Python mirror "
          f"(pageable input, results copied to numpy) {best_w * 1e3:.0f} ms; numpy check ok", flush=True)
