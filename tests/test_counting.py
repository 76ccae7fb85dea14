"""Consumer side: abundance table of k-min-mer hashes (csrc/s2k_count.cuh, counting.py).
CPU tier: the kernels through the host emulation against a numpy dictionary, and the hash-partitioned exchange over
gloo (world size 2).  GPU tier: a config-2 shaped slab through the hot path, counted on the device, against the
dictionary built from the ORACLE's items."""
import importlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _hashes(rng, n, distinct):
    h = rng.integers(0, distinct, n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    if n > 200:
        h[100] = h[7] = np.uint64(0xFFFFFFFFFFFFFFFF)        # the one hash the table cannot store as a key
    return h


def test_count_table_emulated(S, emu_ctx):
    C = importlib.import_module("rust-seq2kminmers_b200.counting")
    rng = np.random.default_rng(3)
    for n, distinct in ((0, 1), (1, 1), (60000, 5000), (20000, 10**9), (30000, 3)):
        h = _hashes(rng, n, distinct)
        got = C.count_local(emu_ctx, h.ctypes.data, n, id_base=1000)
        want = C.count_reference(h, np.arange(n, dtype=np.uint64) + 1000)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), (n, distinct)
        assert int(got[1].sum()) == n
    # explicit ids
    h = _hashes(rng, 5000, 100)
    ids = rng.permutation(5000).astype(np.uint64) + 77
    got = C.count_local(emu_ctx, h.ctypes.data, 5000, d_id_ptr=ids.ctypes.data)
    assert all(np.array_equal(a, b) for a, b in zip(got, C.count_reference(h, ids)))


def test_partition_by_hash_emulated(S, emu_ctx):
    rng = np.random.default_rng(4)
    h = _hashes(rng, 70000, 20000)
    for parts in (1, 2, 3, 8, 64):
        oh, oi = np.zeros(len(h), np.uint64), np.zeros(len(h), np.uint64)
        c = emu_ctx.count_partition_device(h.ctypes.data, len(h), 5, parts, oh.ctypes.data, oi.ctypes.data)
        assert int(c.sum()) == len(h)
        assert np.array_equal(np.sort(oi), np.arange(len(h), dtype=np.uint64) + 5) and np.array_equal(h[(oi - 5).astype(np.int64)], oh)
        dest = np.array([S.Library(emu_ctx.lib.path).c.s2k_count_part(int(x), parts) for x in oh[::97]])
        bounds = np.cumsum(c)
        assert np.array_equal(dest, np.searchsorted(bounds, np.arange(len(h))[::97], side="right"))
    with pytest.raises(Exception):
        emu_ctx.count_partition_device(h.ctypes.data, len(h), 0, 65, oh.ctypes.data, oi.ctypes.data)


WORKER = r'''
import importlib, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["S2K_ROOT"])
S = importlib.import_module("rust-seq2kminmers_b200")
C = importlib.import_module("rust-seq2kminmers_b200.counting")
sharding = importlib.import_module("rust-seq2kminmers_b200.sharding")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(11)
reads = [np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)] for n in [9000, 4000, 150, 12000, 7000, 3000]]
reads += [reads[0].copy(), reads[3][2000:9000].copy()]                 # repeated sequence content: counts above 1
so = np.zeros(len(reads) + 1, dtype=np.uint64); so[1:] = np.cumsum([len(r) for r in reads])
bases = np.concatenate(reads)
b, o, r0 = sharding.shard(bases, so, world, rank)
ctx = S.Context(0, S.Library(os.path.join(os.environ["S2K_ROOT"], "tests", "emu", "libs2k_emu.so")))
got = ctx.run(b, o, 21, 3, 0.1, S.HashMode.HpcSimd)
per, first = sharding.gather_totals(got.n_items, got.n_minimizers)
mine = np.ascontiguousarray(got.hash, dtype=np.uint64)
h, c, f, _ = C.count_distributed(ctx, mine.ctypes.data, len(mine), first, "cpu")
# the whole job's dictionary, built on every rank from the full batch
whole = ctx.run(bases, so, 21, 3, 0.1, S.HashMode.HpcSimd)
wh, wc, wf = C.count_reference(whole.hash)
sel = np.array([ctx.lib.c.s2k_count_part(int(x), world) == rank for x in wh], dtype=bool)
assert np.array_equal(h, wh[sel]) and np.array_equal(c, wc[sel]) and np.array_equal(f, wf[sel]), rank
assert int(wc.max()) >= 2
tot = torch.tensor([int(c.sum()), len(h)]); dist.all_reduce(tot)
assert int(tot[0]) == whole.n_items and int(tot[1]) == len(wh)
dist.barrier(); dist.destroy_process_group()
print(f"rank {rank} ok: {len(h)} distinct of {len(wh)}, {int(c.sum())} items")
'''


def test_count_two_ranks_gloo(tmp_path):
    subprocess.run([str(ROOT / "tests" / "emu" / "build_emu.sh")], check=True)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, S2K_ROOT=str(ROOT), S2K_EMU_SMS="1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


@pytest.mark.gpu
def test_count_config2_slab_against_oracle_dictionary(S, O, gpu_ctx):
    """20 kb x 1500 reads (config 2 shape, two copies of every read so that counts exceed 1) through the hot path on the
    GPU, counted on the device; the expected table is built from the ORACLE's items."""
    import torch
    C = importlib.import_module("rust-seq2kminmers_b200.counting")
    L, n = 20000, 1500
    bases1 = O.synth(0x5EED0002, 0, L * n)
    bases = np.concatenate([bases1, bases1])
    so = np.arange(2 * n + 1, dtype=np.uint64) * np.uint64(L)
    dev = torch.device("cuda", 0)
    d_b = torch.from_numpy(np.concatenate([bases, np.zeros(16, np.uint8)])).to(dev)
    d_so = torch.from_numpy(so.astype(np.int64)).to(dev)
    res = gpu_ctx.run_device(d_b.data_ptr(), d_so.data_ptr(), 2 * n, len(bases), 31, 5, 0.01, S.HashMode.HpcSimd,
                             no_minimizer_stream=True)
    got = C.count_local(gpu_ctx, res.hash, int(res.n_items), device=dev)
    want_h = np.concatenate([O.kminmers(bases1[i * L:(i + 1) * L], 31, 5, 0.01, O.HPCSIMD)["hash"] for i in range(n)])
    want = C.count_reference(np.concatenate([want_h, want_h]))
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    assert int(got[1].min()) >= 2 and int(got[1].sum()) == int(res.n_items)
    # hash-partitioned buckets reproduce the same table when counted part by part
    n_items = int(res.n_items)
    oh = torch.empty(n_items, dtype=torch.int64, device=dev)
    oi = torch.empty(n_items, dtype=torch.int64, device=dev)
    c = gpu_ctx.count_partition_device(res.hash, n_items, 0, 4, oh.data_ptr(), oi.data_ptr())
    parts, a = [], 0
    for p in range(4):
        parts.append(C.count_local(gpu_ctx, oh.data_ptr() + 8 * a, int(c[p]), d_id_ptr=oi.data_ptr() + 8 * a, device=dev))
        a += int(c[p])
    merged = [np.concatenate([x[i] for x in parts]) for i in range(3)]
    o = np.argsort(merged[0], kind="stable")
    assert all(np.array_equal(m[o], w) for m, w in zip(merged, want))
