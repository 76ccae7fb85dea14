#!/usr/bin/env python
"""Long randomized parity run on the GPU (beyond the 120 draws of tests/test_gpu_parity.py): N draws with a given seed,
full tuples and minimizer streams against the oracle.  Usage (GPU box): python tools/fuzz_gpu.py [N] [seed] [max_len]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Batches, assert_batch_matches_oracle
from parity_cases import fuzz_cases
from oracle import oracle as O
S = importlib.import_module("rust-seq2kminmers_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
max_len = int(sys.argv[3]) if len(sys.argv) > 3 else 300000
ctx = S.Context(0)
B = Batches(seed)
t0 = time.time()
import numpy as np
rng = np.random.default_rng(seed + 1)
n_piped = 0
for i, (bases, so, (l, k, d, mode, var)) in enumerate(fuzz_cases(B, n, max_len)):
    if rng.random() < 0.3:                     # every third draw goes through the slab pipeline (whole reads and pieces)
        ctx.set_slab_bytes(int(rng.integers(20000, 200000)))
        ctx.set_transport(int(rng.integers(1, 9)), float(rng.choice([0.0, 0.5, 0.7, 1.0])))
        n_piped += 1
    else:
        ctx.set_slab_bytes(0)
    got = ctx.run(bases, so, l, k, d, S.HashMode(mode), S.HashVariant(var), want_minimizers=True)
    assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, var)
print(f"{n} draws ({n_piped} through small slabs; seed {seed}, max_len {max_len}) identical to the oracle in {time.time() - t0:.1f} s")
