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
