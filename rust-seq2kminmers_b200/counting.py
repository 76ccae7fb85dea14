"""Consumer side of the k-min-mer stream: abundance of every distinct k-min-mer hash.

rust-mdbg inserts the items of ``KminmersIterator`` into a concurrent map keyed by the k-min-mer hash
(``KminmerHash`` is equal / ordered by ``hash`` alone, src/kminmer.rs:181-203; the reference's own remark about the
"Dashmap level", src/lib.rs:256-258).  Here the map is a hash table in HBM (csrc/s2k_count.cuh) and only the distinct
``(hash, count, first item id)`` triples leave the device.

Several GPUs: the items are partitioned BY HASH -- rank ``s2k_count_part(hash, world)`` counts a hash -- so every rank
buckets its items by destination (``s2k_count_partition_device``) and the buckets are exchanged with ONE all-to-all
(NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests): the first data-path collective of this package.  Item ids
are global (rank r's items follow those of ranks < r, the order of ``sharding.gather_totals``), so ``first`` identifies
the first occurrence across the whole job.
"""
from __future__ import annotations

import numpy as np


def count_local(ctx, d_hash_ptr: int, n_items: int, d_id_ptr: int = 0, id_base: int = 0, stream: int = 0, device=None):
    """One GPU.  Returns (hash u64[], count u32[], first u64[]) as host numpy arrays, sorted by hash."""
    import importlib
    S = importlib.import_module(__package__)
    r = ctx.count_device(d_hash_ptr, n_items, d_id_ptr, id_base, stream)
    n = int(r.n_distinct)
    return _fetch(S, r, n, device)


def _fetch(S, r, n, device):
    if n == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint64)
    if device is None or str(device) == "cpu":          # emulated context: "device" memory is host memory
        import ctypes as C
        h = np.ctypeslib.as_array(C.cast(r.hash, C.POINTER(C.c_uint64)), (n,)).copy()
        c = np.ctypeslib.as_array(C.cast(r.count, C.POINTER(C.c_uint32)), (n,)).copy()
        f = np.ctypeslib.as_array(C.cast(r.first, C.POINTER(C.c_uint64)), (n,)).copy()
    else:
        import torch
        h = torch.as_tensor(S.DeviceArray(r.hash, n * 8, "|u1"), device=device).cpu().numpy().view(np.uint64)
        c = torch.as_tensor(S.DeviceArray(r.count, n * 4, "|u1"), device=device).cpu().numpy().view(np.uint32)
        f = torch.as_tensor(S.DeviceArray(r.first, n * 8, "|u1"), device=device).cpu().numpy().view(np.uint64)
    o = np.argsort(h, kind="stable")
    return h[o], c[o], f[o]


def count_distributed(ctx, d_hash_ptr: int, n_items: int, id_base: int, device, group=None, stream: int = 0, fetch: bool = True):
    """All ranks of the default (or given) process group call this with their own items.  Every distinct hash ends up on
    exactly one rank; returns this rank's (hash, count, first) sorted by hash, plus the seconds spent in the exchange.
    fetch=False: the table stays on the device -- returns (s2k_count_result, None, None, seconds)."""
    import importlib
    import time
    import torch
    import torch.distributed as dist
    S = importlib.import_module(__package__)
    world = dist.get_world_size(group)
    dev = torch.device(device)
    send_h = torch.empty(max(n_items, 1), dtype=torch.int64, device=dev)
    send_i = torch.empty(max(n_items, 1), dtype=torch.int64, device=dev)
    counts = ctx.count_partition_device(d_hash_ptr, n_items, id_base, world, send_h.data_ptr(), send_i.data_ptr(), stream)
    send_n = torch.tensor(counts.astype(np.int64), dtype=torch.int64, device=dev)
    recv_n = torch.empty_like(send_n)
    t0 = time.perf_counter()
    dist.all_to_all_single(recv_n, send_n, group=group)
    in_split, out_split = [int(x) for x in send_n.tolist()], [int(x) for x in recv_n.tolist()]
    n_recv = sum(out_split)
    recv_h = torch.empty(max(n_recv, 1), dtype=torch.int64, device=dev)
    recv_i = torch.empty(max(n_recv, 1), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_h[:n_recv], send_h[:n_items], out_split, in_split, group=group)
    dist.all_to_all_single(recv_i[:n_recv], send_i[:n_items], out_split, in_split, group=group)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    t_exchange = time.perf_counter() - t0
    r = ctx.count_device(recv_h.data_ptr(), n_recv, recv_i.data_ptr(), 0, stream)
    if not fetch:
        return r, None, None, t_exchange
    h, c, f = _fetch(S, r, int(r.n_distinct), None if dev.type == "cpu" else dev)
    return h, c, f, t_exchange


def count_reference(hashes, ids=None):
    """The same table on the host with numpy (tests: the oracle-side dictionary)."""
    hashes = np.asarray(hashes, dtype=np.uint64)
    ids = np.arange(len(hashes), dtype=np.uint64) if ids is None else np.asarray(ids, dtype=np.uint64)
    if len(hashes) == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint64)
    o = np.lexsort((ids, hashes))
    hs, is_ = hashes[o], ids[o]
    firsts = np.concatenate([[True], hs[1:] != hs[:-1]])
    idx = np.flatnonzero(firsts)
    cnt = np.diff(np.concatenate([idx, [len(hs)]])).astype(np.uint32)
    return hs[idx], cnt, is_[idx]
