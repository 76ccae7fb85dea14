"""Multi-GPU partitioning for the sequence -> k-min-mer path.

Sequences are independent (the reference fans records out to host threads, src/main.rs:65-79), so the batch is
cut into contiguous read ranges balanced by base count, one range per rank / GPU.  Each rank runs its range
through its own Context; no collective sits on the data path.  The only collective is the optional gather of
per-rank totals (items, minimizers), from which a rank derives the global index of its first item.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def partition_reads(seq_off, world: int) -> List[Tuple[int, int]]:
    """Contiguous read ranges [r0, r1) per rank, balanced by bases; every read lands in exactly one range."""
    so = np.asarray(seq_off, dtype=np.uint64)
    n = len(so) - 1
    total = int(so[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        cut = int(np.searchsorted(so, np.uint64(target), side="left"))
        cuts.append(min(max(cut, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def shard(bases, seq_off, world: int, rank: int):
    """(bases slice, rebased offsets, first read index) of `rank`'s shard."""
    so = np.asarray(seq_off, dtype=np.uint64)
    r0, r1 = partition_reads(so, world)[rank]
    b0, b1 = int(so[r0]), int(so[r1])
    return np.asarray(bases)[b0:b1], so[r0:r1 + 1] - so[r0], r0


def gather_totals(n_items: int, n_minimizers: int, device=None):
    """all_gather of (items, minimizers) over the default process group (NCCL on GPUs, gloo on CPU).
    Returns (per_rank int64[world, 2], first global item index of this rank)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([n_items, n_minimizers], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    per = torch.stack(out).cpu().numpy()
    return per, int(per[:rank, 0].sum())
