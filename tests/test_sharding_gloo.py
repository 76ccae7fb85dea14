"""CPU tier, world_size 2 over gloo: the multi-GPU host logic (read partitioning by bases, per-rank runs, gather of
totals).  Each rank drives the kernels through the test-tier host emulation (tests/emu); on a GPU box the same
code path runs with one Context per device and NCCL (bench.py --gpus N)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import importlib, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["S2K_ROOT"])
S = importlib.import_module("rust-seq2kminmers_b200")
sharding = importlib.import_module("rust-seq2kminmers_b200.sharding")
from oracle import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)
lens = [20000, 150, 0, 9000, 150, 150, 31, 12000, 700, 5000]
seqs = [np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)] for n in lens]
so = np.zeros(len(lens) + 1, dtype=np.uint64); so[1:] = np.cumsum(lens)
bases = np.concatenate(seqs)
parts = sharding.partition_reads(so, world)
assert parts[0][0] == 0 and parts[-1][1] == len(lens) and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
b, o, r0 = sharding.shard(bases, so, world, rank)
ctx = S.Context(0, S.Library(os.path.join(os.environ["S2K_ROOT"], "tests", "emu", "libs2k_emu.so")))
got = ctx.run(b, o, 31, 5, 0.02, S.HashMode.HpcSimd)
per, first = sharding.gather_totals(got.n_items, got.n_minimizers)
want = [O.kminmers(s, 31, 5, 0.02, O.HPCSIMD) for s in seqs]
assert int(per[:, 0].sum()) == sum(len(w["hash"]) for w in want)
assert first == sum(len(w["hash"]) for w in want[:r0])
for i in range(len(o) - 1):
    a, e = int(got.km_off[i]), int(got.km_off[i + 1])
    w = want[r0 + i]
    assert np.array_equal(got.hash[a:e], w["hash"]) and np.array_equal(got.start[a:e], w["start"].astype(np.uint32))
    assert np.array_equal(got.end[a:e], w["end"].astype(np.uint32)) and np.array_equal(got.rev[a:e], w["rev"])
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank} ok: reads [{parts[rank][0]},{parts[rank][1]}) items {got.n_items}")
'''


def test_partition_is_balanced_and_complete():
    sys.path.insert(0, str(ROOT))
    import importlib
    sharding = importlib.import_module("rust-seq2kminmers_b200.sharding")
    so = np.concatenate([[0], np.cumsum(np.random.default_rng(1).integers(0, 30000, 1000))]).astype(np.uint64)
    for world in (1, 2, 3, 8):
        parts = sharding.partition_reads(so, world)
        assert parts[0][0] == 0 and parts[-1][1] == 1000
        sizes = [int(so[b] - so[a]) for a, b in parts]
        assert sum(sizes) == int(so[-1]) and max(sizes) - min(sizes) <= 2 * 30000
    assert sharding.partition_reads(np.array([0], dtype=np.uint64), 4) == [(0, 0)] * 4


def test_two_ranks_gloo(tmp_path):
    subprocess.run([str(ROOT / "tests" / "emu" / "build_emu.sh")], check=True)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, S2K_ROOT=str(ROOT), S2K_EMU_SMS="1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_one_sequence_split_across_ranks_emulated():
    """one long sequence cut into base ranges; overlap-and-trim by ownership == the whole sequence."""
    import importlib
    sys.path.insert(0, str(ROOT))
    S = importlib.import_module("rust-seq2kminmers_b200")
    sharding = importlib.import_module("rust-seq2kminmers_b200.sharding")
    from oracle import oracle as O
    subprocess.run([str(ROOT / "tests" / "emu" / "build_emu.sh")], check=True)
    ctx = S.Context(0, S.Library(ROOT / "tests" / "emu" / "libs2k_emu.so"))
    rng = np.random.default_rng(11)
    base = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 90000)]
    base[30000:30200] = ord("A")                      # a homopolymer across a cut
    l, k, d = 31, 5, 0.02
    for mode, var in [(1, 0), (3, 0), (2, 0), (0, 0), (3, 1)]:
        seqs = [base]
        if (mode, var) == (3, 0):                      # make the AVX-512 tail rule fire: trim until S % 16 == 0
            for cut in range(0, 64):
                s = base[:len(base) - cut]
                if (len(O.encode_rle_simd(s)[0]) - l + 1) % 16 == 0:
                    seqs.append(s)
                    break
        for seq in seqs:
            want = O.kminmers(seq, l, k, d, mode, var)
            for world in (2, 3):
                ranges = sharding.sequence_ranges(len(seq), world, sharding.default_overlap_right(l, k, d))
                kept = [sharding.kept_in_range(seq[lo:hi], lo, b0, b1, mode in (1, 3)) for (b0, b1, lo, hi) in ranges]
                parts = []
                for r, (b0, b1, lo, hi) in enumerate(ranges):
                    parts.append(sharding.run_sequence_part(ctx, seq[lo:hi], lo, b0, b1, len(seq), l, k, d, S.HashMode(mode),
                                                            S.HashVariant(var), kept_total=sum(kept)))
                for key in ("hash", "start", "end", "rev"):
                    cat = np.concatenate([p[key] for p in parts])
                    assert np.array_equal(cat, want[key].astype(cat.dtype)), (mode, var, world, key, len(seq))
    ctx.close()
