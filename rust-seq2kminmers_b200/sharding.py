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


# ---------------------------------------------------------------------------------------------- one long sequence
# a single sequence (a chromosome) is split into contiguous base ranges, one per GPU.  No exchange of
# bases or minimizers is needed: rank r processes its range [b0, b1) extended by a small left overlap and a right
# overlap as if it were a sequence of its own, and keeps what it OWNS -- the minimizers whose start lies in [b0, b1) and
# the k-min-mers whose first minimizer it owns.  Every l-mer that starts after the first base of the extended range
# sees the same bases as in the whole sequence, so owned minimizers are identical to the global ones; a window that
# begins in [b0, b1) is completed with minimizers found in the right overlap (recomputed, not received).  Two things
# need the whole sequence and travel as scalars over the process group: the number of windows before the rank (the
# `offset` base) and, for the AVX-512 tail rule of ntHash1 (src/nthash_avx512_32.rs:134-138), the HPC length.

OVERLAP_LEFT = 64


def sequence_ranges(n_bases: int, world: int, overlap_right: int):
    """[(b0, b1, lo, hi)] per rank: owned range and the extended range that has to be resident on the rank."""
    out = []
    for r in range(world):
        b0, b1 = n_bases * r // world, n_bases * (r + 1) // world
        out.append((b0, b1, max(0, b0 - OVERLAP_LEFT), min(n_bases, b1 + overlap_right)))
    return out


def default_overlap_right(l: int, k: int, density: float) -> int:
    """Raw bases that hold k-1 further minimizers with a wide margin (selection rate >= density per kept base)."""
    rate = max(min(density, 1.0), 1e-9)
    return int(min(2 ** 31, 64 * (k + 8) / rate + 64 * l + 4096))


def kept_in_range(part: np.ndarray, lo: int, b0: int, b1: int, hpc: bool) -> int:
    """Kept (HPC) bases of the whole sequence that fall into [b0, b1); `part` holds bases [lo, ...)."""
    if not hpc:
        return b1 - b0
    seg = np.asarray(part)[b0 - lo:b1 - lo]
    if len(seg) == 0:
        return 0
    prev = np.asarray(part)[b0 - lo - 1:b1 - lo - 1] if b0 > 0 else None
    if prev is None:
        return 1 + int(np.count_nonzero(seg[1:] != seg[:-1]))
    return int(np.count_nonzero(seg != prev))


def run_sequence_part(ctx, part, lo: int, b0: int, b1: int, n_total: int, l: int, k: int, density: float, mode, variant=0,
                      kept_total: int = None):
    """k-min-mers of ONE sequence owned by the range [b0, b1); `part` = bases [lo, hi) of it, resident on this rank.
    Returns dict(hash u64, start u64, end u64, rev u8, n_minimizers) in global coordinates, in sequence order.
    kept_total (HPC length of the whole sequence = sum over ranks of kept_in_range, one all_gather) is needed only in
    the ntHash1 Simd/HpcSimd modes, by ranks whose extended range reaches the end of the sequence (tail rule)."""
    part = np.ascontiguousarray(part, dtype=np.uint8)
    hi = lo + len(part)
    is_last = b1 == n_total
    mode_i, var_i = int(mode), int(variant)
    simd_nt1 = mode_i in (2, 3) and var_i == 0
    if n_total <= l:                                   # src/lib.rs:97
        e = np.empty(0, np.uint64)
        return dict(hash=e, start=e, end=e, rev=np.empty(0, np.uint8), n_minimizers=0)
    got = ctx.run(part, np.array([0, len(part)], dtype=np.uint64), l, k, density, mode, variant, want_minimizers=True,
                  no_tail_rule=True)
    mins = got.minimizers
    n_min = len(mins)
    if simd_nt1 and hi == n_total:                     # the rule is a property of the whole sequence
        hpc = mode_i == 3
        if kept_total is None:
            raise ValueError("ranks that reach the end of the sequence need kept_total in the ntHash1 Simd/HpcSimd modes")
        S = kept_total - l + 1
        if S > 16 and S % 16 == 0:                     # drop the minimizers whose last base is among the final 16 kept bases
            if hpc:
                tail = part[-min(len(part), 1 << 20):]
                keep = np.flatnonzero(np.concatenate([[True], tail[1:] != tail[:-1]]))
                if len(keep) < 17 and len(tail) < len(part):
                    raise ValueError("homopolymer tail longer than the inspected window")
                e16 = hi - len(tail) + int(keep[-16])
            else:
                e16 = n_total - 16
            while n_min > 0 and lo + int(mins["end"][n_min - 1]) >= e16:
                n_min -= 1
    starts = mins["start"][:n_min].astype(np.uint64) + np.uint64(lo)
    i0 = int(np.searchsorted(starts, np.uint64(b0), side="left"))
    i1 = n_min if is_last else int(np.searchsorted(starts, np.uint64(b1), side="left"))
    n_items = max(0, n_min - k + 1)
    if not is_last and hi < n_total and i1 - 1 + (k - 1) + 40 >= n_min:
        raise ValueError("right overlap too short: fewer than k-1 minimizers (plus margin) after the owned range")
    w1 = min(i1, n_items)
    w0 = min(i0, w1)
    sl = slice(w0, w1)
    return dict(hash=got.hash[sl].copy(), start=got.start[sl].astype(np.uint64) + np.uint64(lo),
                end=got.end[sl].astype(np.uint64) + np.uint64(lo), rev=got.rev[sl].copy(), n_minimizers=i1 - i0)
