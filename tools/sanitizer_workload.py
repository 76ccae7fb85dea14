import sys, numpy as np
sys.path.insert(0, '.')
import seq2kminmers_b200 as S
from oracle import oracle as O
rng = np.random.default_rng(5)
lens = [0, 31, 32, 150, 150, 9000, 0, 20000, 47, 48, 40000, 3]
seqs = [np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 4 if i % 4 else 5, n)] for i, n in enumerate(lens)]
seqs.append(np.full(30000, ord('A'), np.uint8)); seqs.append(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 5000)])
so = np.zeros(len(seqs) + 1, dtype=np.uint64); so[1:] = np.cumsum([len(s) for s in seqs]); bases = np.concatenate(seqs)
ctx = S.Context(0)
for mode, var in [(3, 0), (1, 0), (2, 0), (0, 0), (3, 1)]:
    r = ctx.run(bases, so, 31, 5, 0.02, S.HashMode(mode), S.HashVariant(var), want_minimizers=True)
    for i, s in enumerate(seqs):
        w = O.kminmers(s, 31, 5, 0.02, mode, var)
        a, b = int(r.km_off[i]), int(r.km_off[i + 1])
        assert np.array_equal(r.hash[a:b], w["hash"]), (mode, var, i)
ctx.set_slab_bytes(30000)
r = ctx.run(bases, so, 31, 5, 0.02, S.HashMode.HpcSimd)
h, p, off = ctx.encode_rle(bases, so)
print("sanitizer workload ok", r.n_items, len(h))
