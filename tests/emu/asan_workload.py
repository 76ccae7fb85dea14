"""Adversarial batch through the emulated kernels built with AddressSanitizer (see tests/test_emu_asan.py)."""
import sys, numpy as np
import os
sys.path.insert(0, os.environ['S2K_ROOT'])
import seq2kminmers_b200 as S
rng = np.random.default_rng(5)
lens = [0, 31, 32, 150, 150, 9000, 0, 20000, 47, 48, 40000, 3, 255, 256, 700]
seqs = [np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 4 if i % 4 else 5, n)] for i, n in enumerate(lens)]
seqs.append(np.full(30000, ord('A'), np.uint8)); seqs.append(np.frombuffer(b"ACGT" * 9000, dtype=np.uint8).copy())
so = np.zeros(len(seqs) + 1, dtype=np.uint64); so[1:] = np.cumsum([len(s) for s in seqs]); bases = np.concatenate(seqs)
# exact-size buffers so that any over-read of the inputs is caught
bases = bases.copy(); so = so.copy()
ctx = S.Context(0, S.Library(os.environ['S2K_ASAN_LIB']))
for mode, var, l in [(3, 0, 31), (1, 0, 31), (0, 0, 31), (3, 1, 31), (1, 0, 255), (3, 0, 1), (1, 2, 40), (1, 3, 31), (0, 3, 20)]:
    r = ctx.run(bases, so, l, 5, 0.05, S.HashMode(mode), S.HashVariant(var), want_minimizers=True)
    print("ok", mode, var, l, r.n_items)
ctx.set_slab_bytes(30000)
r = ctx.run(bases, so, 31, 5, 0.02, S.HashMode.HpcSimd, want_minimizers=True)
h, p, off = ctx.encode_rle(bases, so)
# device API with S2K_NO_MINIMIZER_STREAM: the window stage and the tail rule read the records in place (tile index walks)
import ctypes
so_dev = so.copy()
for mode, k, d in [(3, 5, 0.05), (0, 12, 0.002)]:
    want = ctx.run(bases, so, 31, k, d, S.HashMode(mode))
    rd = ctx.run_device(bases.ctypes.data, so_dev.ctypes.data, len(so) - 1, len(bases), 31, k, d, S.HashMode(mode),
                        no_minimizer_stream=True)
    assert not rd.minimizers and int(rd.n_items) == want.n_items, (mode, k, d)
    got = np.frombuffer((ctypes.c_uint64 * int(rd.n_items)).from_address(rd.hash), dtype=np.uint64) if rd.n_items else np.zeros(0, np.uint64)
    assert np.array_equal(got, want.hash), (mode, k, d)
    print("ok in place", mode, k, d, int(rd.n_items))
print("asan workload done", r.n_items, len(h))
