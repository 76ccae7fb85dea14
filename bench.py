#!/usr/bin/env python
"""bench.py -- input Gbp/s -> k-min-mers on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c4off|c5]

A step = one pass of the whole hot path (HPC -> ntHash -> threshold -> ordered minimizers -> k-window hash)
over one batch of synthetic reads.  N=1 workload = BASELINE.json configs[1]: 500 000 HiFi-like reads x 20 kb
(10 Gbp), HPC on, ntHash1, l=31 k=5 density=0.01, HashMode::HpcSimd.  Multi-GPU: reads are sharded by rank,
no collective on the data path (weak scaling: each rank processes a full per-GPU workload).

  value    : device-resident throughput, inputs already in HBM, CUDA events, max over ranks
  e2e      : same metric through s2k_run with pinned HOST buffers (H2D of bases+offsets and D2H of all items
             inside the timed region); e2e.roofline = what the host side can feed (probed H2D and packing rates)
  roofline : HBM roofline of the dominant kernel (k_minimizers): algorithmic bytes / its CUDA-event time
  parity   : AFTER the timed region every read's order-sensitive digest of its full (hash, start, end, rev)
             tuples is compared with the CPU oracle's (oracle/ is the checker, never the measured path)
  extra    : the other single-GPU BASELINE shapes in the same run (configs[2] 150-bp reads, configs[3] one
             3.1-Gbp sequence), split over the ranks (strong scaling), each with its own roofline and parity
  cpu_baseline / --impl reference : the CPU oracle port of the reference's iterator on all host cores
             (the Rust reference itself cannot be built in this image: no cargo, nightly-only, git deps)
  --workload c5 : BASELINE configs[4], the 30-Gbp sweep over density x k x HPC, with a CPU-port column
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (read_len, n_reads, seed, mode, variant, description)
    "c2": (20000, 500_000, 0x5EED0002, 3, 0, "configs[1]: synthetic HiFi-like reads 20 kb x 500k (10 Gbp), HPC on, ntHash1, HashMode::HpcSimd"),
    "c3": (150, 100_000_000, 0x5EED0003, 3, 0, "configs[2]: synthetic 150 bp reads x 100M (15 Gbp), HPC on, ntHash1, HashMode::HpcSimd"),
    "c4": (3_100_000_000, 1, 0x5EED0004, 3, 1, "configs[3]: one 3.1 Gbp sequence, chunked with halos, ntHash2-31, HPC on"),
    "c4off": (3_100_000_000, 1, 0x5EED0004, 2, 1, "configs[3]: one 3.1 Gbp sequence, chunked with halos, ntHash2-31, HPC off (HashMode::Simd)"),
}
L_PARAM, K_PARAM, DENSITY = 31, 5, 0.01
MODE_NAMES = ["Regular", "Hpc", "Simd", "HpcSimd"]


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_port_rate(O, L, seed, mode, variant, threads, target_s=12.0, first_read=0, l=L_PARAM, k=K_PARAM, density=DENSITY):
    """Times the oracle port (mirrors src/main.rs:65-79: one iterator per read, count items) on a bounded sample."""
    probe_reads = max(threads * 4, int(2_000_000 // max(L, 1)) + 1) if L < 10_000_000 else 1
    probe_len = L if L < 10_000_000 else 20_000_000

    use_avx = variant == 0 and mode in (2, 3) and O.has_avx512()

    def run(n_reads, read_len):
        bases = O.synth(seed, first_read * L, n_reads * read_len)
        so = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
        t0 = time.perf_counter()
        if use_avx:      # AVX-512 restatement of the reference's vector path (oracle/s2k_cpu_avx512.c)
            r = O.avx512_batch(bases, so, l, k, density, mode, threads=threads, want_counts=False)
        else:            # scalar restatement (oracle/s2k_oracle.c)
            r = O.batch(bases, so, l, k, density, mode, variant, threads=threads, want_counts=False)
        return time.perf_counter() - t0, n_reads * read_len, r["total"]

    # Bounded sample: one slab of at most ~2 Gbp, repeated until about target_s seconds of CPU work have been timed.
    dt, nb, _ = run(probe_reads, probe_len)
    rate = nb / dt
    n_reads = max(probe_reads, min(int(rate * target_s / probe_len), int(2e9 // probe_len) or 1))
    bases = O.synth(seed, first_read * L, n_reads * probe_len)
    so = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(probe_len)
    call = (lambda: O.avx512_batch(bases, so, l, k, density, mode, threads=threads, want_counts=False)) if use_avx \
        else (lambda: O.batch(bases, so, l, k, density, mode, variant, threads=threads, want_counts=False))
    call()                                             # warm the threads/caches
    dt, reps = 0.0, 0
    while dt < target_s and reps < 64:
        t0 = time.perf_counter()
        call()
        dt += time.perf_counter() - t0
        reps += 1
    nb = n_reads * probe_len * reps
    impl = "AVX-512 C restatement of the reference's vector path (oracle/s2k_cpu_avx512.c)" if use_avx else \
        "scalar C restatement of the reference iterator (oracle/s2k_oracle.c)"
    return nb / dt / 1e9, dt, f"{n_reads} reads x {probe_len} bp x {reps} repeats = {nb / 1e9:.3f} Gbp of the same synthetic stream, {dt:.1f} s; {impl}"


def workload_config(name, world, reads_override=0, mode_override=-1):
    L, n_reads, seed, mode, variant, desc = WORKLOADS[name]
    if reads_override:
        n_reads = reads_override
    if mode_override >= 0:
        mode = mode_override
    return {"workload": desc, "read_len": L, "reads_per_gpu": n_reads, "l": L_PARAM, "k": K_PARAM, "density": DENSITY,
            "hash_mode": MODE_NAMES[mode], "hash": "ntHash2-31" if variant else "ntHash1-32",
            "sharding": f"reads split by rank x{world}, no data-path collective",
            "l2": "inputs (>=3 GB per step) far exceed the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.S = importlib.import_module("rust-seq2kminmers_b200")
        self.sharding = importlib.import_module("rust-seq2kminmers_b200.sharding")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = self.S.Context(self.local_rank)
        self.ctx.set_timing(True)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.threads = os.cpu_count() or 1
        self.peak, self.peak_src = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allmin(self, x):
        return -self.allmax(-x)

    def allsum(self, xs):
        t = self.torch.tensor(list(xs), dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(v) for v in t.tolist()]

    # ------------------------------------------------------------------ device-resident inputs of one workload
    def setup(self, name, scaling, reads_override=0, mode_override=-1, seed_override=None, k=K_PARAM, density=DENSITY):
        """scaling 'weak': every rank holds the whole per-GPU workload (its own slice of the synthetic stream);
        'strong': the workload is split over the ranks (reads by rank; one long sequence into base ranges)."""
        torch, S = self.torch, self.S
        L, n_reads, seed, mode, variant, desc = WORKLOADS[name]
        if reads_override:
            n_reads = reads_override
        if mode_override >= 0:
            mode = mode_override
        if seed_override is not None:
            seed = seed_override
        W = {"name": name, "L": L, "seed": seed, "mode": mode, "variant": variant, "desc": desc, "k": k, "density": density,
             "scaling": scaling, "split_one": False, "lo": 0}
        if n_reads == 1 and self.world > 1 and scaling == "strong":
            b0, b1, lo, hi = self.sharding.sequence_ranges(L, self.world, self.sharding.default_overlap_right(L_PARAM, k, density))[self.rank]
            W.update(split_one=True, b0=b0, b1=b1, lo=lo, hi=hi, n_reads=1, n_bases=hi - lo, first_base=lo, total_bases=L)
        elif scaling == "strong" and self.world > 1:
            r0, r1 = n_reads * self.rank // self.world, n_reads * (self.rank + 1) // self.world
            W.update(n_reads=r1 - r0, n_bases=(r1 - r0) * L, first_base=r0 * L, total_bases=n_reads * L)
        else:
            W.update(n_reads=n_reads, n_bases=n_reads * L, first_base=self.rank * n_reads * L, total_bases=self.world * n_reads * L)
        W["d_bases"] = torch.empty(W["n_bases"] + 16, dtype=torch.uint8, device=self.dev)
        self.ctx.synth_device(seed, W["first_base"], W["n_bases"], W["d_bases"].data_ptr())
        if W["n_reads"] == 1:
            W["d_so"] = torch.tensor([0, W["n_bases"]], dtype=torch.int64, device=self.dev)
        else:
            W["d_so"] = torch.arange(W["n_reads"] + 1, dtype=torch.int64, device=self.dev) * L
        torch.cuda.synchronize()
        W["in_place"] = self.args.minimizer_stream == "in-place" and not W["split_one"]   # a rank of a split sequence needs the stream
        return W

    def step(self, W):
        S = self.S
        return self.ctx.run_device(W["d_bases"].data_ptr(), W["d_so"].data_ptr(), W["n_reads"], W["n_bases"], L_PARAM, W["k"],
                                   W["density"], S.HashMode(W["mode"]), S.HashVariant(W["variant"]), stream=self.stream,
                                   no_tail_rule=W["split_one"], no_minimizer_stream=W["in_place"])

    def owned(self, W, res):
        """(items, minimizers) this rank contributes: everything, or for a split sequence what it OWNS."""
        torch, S = self.torch, self.S
        n_items, n_min = int(res.n_items), int(res.n_minimizers)
        if W["split_one"]:
            mins_t = torch.as_tensor(S.DeviceArray(res.minimizers, n_min * 16, "|u1"), device=self.dev).view(torch.int32).view(-1, 4)
            starts = (mins_t[:, 1].to(torch.int64) & 0xffffffff) + W["lo"]
            i0 = int(torch.searchsorted(starts, torch.tensor([W["b0"]], device=self.dev))[0])
            i1 = n_min if self.rank == self.world - 1 else int(torch.searchsorted(starts, torch.tensor([W["b1"]], device=self.dev))[0])
            n_items, n_min = max(0, min(i1, max(0, n_min - W["k"] + 1)) - i0), i1 - i0
        return n_items, n_min

    # ------------------------------------------------------------------ device-resident timing
    def time_resident(self, W, steps, warmup, sample_clocks=False):
        torch = self.torch
        for _ in range(max(warmup, 3)):
            res = self.step(W)
        n_items, n_min = self.owned(W, res)
        self.barrier()
        sampler = ClockSampler(self.local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        launches0 = self.ctx.launch_count
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        kms, wms, kl = [], [], 0
        ev[0].record()
        for _ in range(steps):
            res = self.step(W)
            a, b, c = self.ctx.last_kernel_ms()
            kms.append(a); wms.append(b); kl = c
        ev[1].record()
        self.barrier()
        launches = self.ctx.launch_count - launches0
        ms_step = self.allmax(ev[0].elapsed_time(ev[1])) / steps
        clocks = sampler.summary() if sampler else None
        value = W["total_bases"] / (ms_step * 1e-3) / 1e9
        tot_items, tot_min = self.allsum([n_items, n_min])
        # roofline of the dominant kernel (k_minimizers): algorithmic bytes = bases + offsets + 17 B per item
        alg_bytes = W["n_bases"] + 8 * (W["n_reads"] + 1) + 17 * n_items
        k_ms = float(np.mean(kms))
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak,
                    "traffic": None, "traffic_source": None, "algorithmic_bytes": alg_bytes, "kernel": "k_minimizers",
                    "launches_per_step": kl, "ms_per_step_in_kernel": k_ms, "window_stage_ms": float(np.mean(wms)),
                    "bytes_per_base": alg_bytes / max(1, W["n_bases"]), "peak_source": self.peak_src}
        return {"value": value, "ms_per_step": ms_step, "roofline": roofline, "clocks": clocks, "launches": int(launches),
                "items": tot_items, "minimizers": tot_min, "res": res, "n_items_rank": n_items}

    # ------------------------------------------------------------------ full-scale parity (outside every timed region)
    def device_items(self, res, n_reads):
        """The result of the last device run, copied to host numpy arrays."""
        torch, S = self.torch, self.S
        n = int(res.n_items)
        def get(ptr, count, ts, dt):
            if count == 0:
                return np.zeros(0, dtype=dt)
            return torch.as_tensor(S.DeviceArray(ptr, count, ts), device=self.dev).cpu().numpy().view(dt)
        return (get(res.hash, n * 8, "|u1", np.uint64), get(res.start, n * 4, "|u1", np.uint32), get(res.end, n * 4, "|u1", np.uint32),
                get(res.rev, n, "|u1", np.uint8), get(res.km_off, (n_reads + 1) * 8, "|u1", np.uint64))

    def parity(self, O, W, res, host_bases=None):
        """Every read of this rank's batch: the order-sensitive digest of its full item tuples (oracle/s2k_oracle.c: fold)
        against the CPU oracle run on the same bases -- the AVX-512 restatement for ntHash1 simd profiles, the scalar
        restatement otherwise.  One long sequence: the oracle runs on overlapping pieces (cut on run boundaries) and every
        item is compared in the piece it is interior to.  Split sequences (c4 on several GPUs) are covered at N=1."""
        if W["split_one"]:
            return {"checked": False, "why": "one sequence split over ranks: full parity runs at N=1 (and in tests/test_gpu_parity.py on 4 ranges)"}
        torch = self.torch
        t0 = time.perf_counter()
        hb = host_bases if host_bases is not None else W["d_bases"][:W["n_bases"]].cpu().numpy()
        h, s, e, rv, km = self.device_items(res, W["n_reads"])
        mode, var, k, d = W["mode"], W["variant"], W["k"], W["density"]
        if W["n_reads"] == 1:
            ok, checked, n_pieces = long_sequence_parity(O, hb, h, s, e, rv, L_PARAM, k, d, mode, var, self.threads)
            out = {"checked": True, "reads_checked": 1, "items_checked": int(checked), "items": int(len(h)), "pieces": n_pieces,
                   "digest_match": bool(ok and checked == len(h)),
                   "how": "oracle on overlapping pieces cut on run boundaries; every item compared field by field in the piece it is interior to"}
        else:
            so = np.arange(W["n_reads"] + 1, dtype=np.uint64) * np.uint64(W["L"])
            use_avx = var == 0 and mode in (2, 3) and O.has_avx512()
            if use_avx:
                want = O.avx512_batch(hb, so, L_PARAM, k, d, mode, threads=self.threads, want_digest=True)
            else:
                want = O.batch(hb, so, L_PARAM, k, d, mode, var, threads=self.threads, want_digest=True)
            got = O.digest_items(h, s, e, rv, km)
            cnt_ok = bool(np.array_equal(np.diff(km), want["km_cnt"]))
            dg_ok = bool(np.array_equal(got, want["digest"]))
            out = {"checked": True, "reads_checked": int(W["n_reads"]), "items": int(len(h)), "digest_match": bool(cnt_ok and dg_ok),
                   "counts_match": cnt_ok, "how": ("AVX-512" if use_avx else "scalar") + " CPU oracle, per-read order-sensitive digest of (hash, start, end, rev)"}
        out["seconds"] = round(time.perf_counter() - t0, 2)
        # every rank checks its own reads; the line reports the conjunction and the total
        ok_all, reads_all = self.allsum([1 if out["digest_match"] else 0, out["reads_checked"]])
        out["digest_match"] = ok_all == self.world
        out["reads_checked"] = reads_all
        return out

    # ------------------------------------------------------------------ host probes: what can feed the GPUs
    def host_probes(self, hb_np, hb_t):
        """All ranks at once (that is how they share the host): pinned H2D rate of this rank's link, and the rate at which
        this rank's share of the host threads packs bases to 2 bits."""
        torch = self.torch
        n = min(len(hb_np), 2_000_000_000)
        d = torch.empty(n, dtype=torch.uint8, device=self.dev)
        self.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        d.copy_(hb_t[:n], non_blocking=True)
        torch.cuda.synchronize()
        self.barrier()
        ev[0].record()
        for _ in range(2):
            d.copy_(hb_t[:n], non_blocking=True)
        ev[1].record()
        torch.cuda.synchronize()
        h2d = 2 * n / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
        del d
        pack_threads = max(1, min(16, (self.threads * 3 // 4) // self.world))
        m = min(n, 1_000_000_000)
        self.ctx.pack2(hb_np[:m // 8], pack_threads)
        self.barrier()
        t0 = time.perf_counter()
        self.ctx.pack2(hb_np[:m], pack_threads)
        pack = m / (time.perf_counter() - t0) / 1e9
        # ... and the two at the same time, results flowing back as well: that is how s2k_run uses the host (the packers,
        # the H2D DMA reads and the D2H DMA writes share the host's memory channels), so this is the tighter ceiling
        d = torch.empty(n, dtype=torch.uint8, device=self.dev)
        back = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
        s2 = torch.cuda.Stream(device=self.dev)
        k_copies = 4
        reps = max(1, int(round(k_copies * n / (h2d * 1e9) * pack * 1e9 / m)))
        self.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(k_copies):
            d.copy_(hb_t[:n], non_blocking=True)
        ev[1].record()
        with torch.cuda.stream(s2):
            for _ in range(k_copies):
                back.copy_(d[:n // 4], non_blocking=True)
        t0 = time.perf_counter()
        for _ in range(reps):
            self.ctx.pack2(hb_np[:m], pack_threads)
        t_pack = time.perf_counter() - t0
        torch.cuda.synchronize()
        h2d_c = k_copies * n / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
        pack_c = reps * m / t_pack / 1e9
        del d, back
        return {"h2d_gbs": h2d, "pack_gbps": pack, "pack_threads": pack_threads, "ranks_probing_together": self.world,
                "concurrent": {"h2d_gbs": h2d_c, "pack_gbps": pack_c,
                               "what": "H2D copies, D2H copies of a quarter of the bytes and the packers running together"}}

    @staticmethod
    def choose_transport(pr):
        """Hybrid transport: a fraction f of the slabs is packed (0.25 B/base over PCIe, costs host threads), the rest goes
        as ASCII.  PCIe time per base (1 - 0.75 f) / R and packing time f / P run side by side: f* = 1 / (0.75 + R / P).
        Packing is kept only if the predicted gain over plain ASCII exceeds 10 % (it also costs host memory bandwidth
        that the other ranks' DMA engines want)."""
        R, P = pr["h2d_gbs"], pr["pack_gbps"]
        f = 1.0 / (0.75 + R / max(P, 1e-9))
        f = max(0.0, min(1.0, f))
        predicted = R / (1.0 - 0.75 * f)
        if predicted < 1.10 * R:
            return 0, 0.0, R
        return pr["pack_threads"], round(f, 2), predicted

    def e2e(self, W, hb_np, hso_np, n_items, probes, pack_threads, pack_ratio, bound):
        torch, S, ctx = self.torch, self.S, self.ctx
        mode, var, k, d = W["mode"], W["variant"], W["k"], W["density"]
        ctx.set_transport(pack_threads, pack_ratio)
        run = lambda **kw: ctx.run(hb_np, hso_np, L_PARAM, k, d, S.HashMode(mode), S.HashVariant(var), copy=False,
                                   no_tail_rule=W["split_one"], **kw)
        e_steps = self.args.steps
        out = run()                                        # warm-up
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            out = run()
        torch.cuda.synchronize()
        dt = self.allmax(time.perf_counter() - t0)
        assert W["split_one"] or out.n_items == n_items
        h2d_actual, n_packed, n_plain = ctx.last_transport()
        value = W["total_bases"] * e_steps / dt / 1e9
        e2e = {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": int(h2d_actual),
               "host_input_bytes_per_step": W["n_bases"] + 8 * (W["n_reads"] + 1),
               "d2h_bytes_per_step": 17 * int(out.n_items) + 8 * (W["n_reads"] + 1) * 2 + 4 * W["n_reads"],
               "steps": e_steps, "api": "s2k_run (C ABI, pinned host ASCII buffers in, pinned host items out)",
               "transport": f"{n_packed} slabs packed to 2 bits/base by host threads + {n_plain} slabs as plain ASCII "
                            f"(s2k_ctx_set_transport: threads {pack_threads or 'default'}, ratio {pack_ratio}; chosen from the probes)",
               "probes": probes,
               "roofline": {"bound": "host feed: min over ranks of the probed H2D rate combined with the probed packing rate (all ranks probing at once)",
                            "peak": self.world * bound, "unit": "Gbp/s",
                            "frac": value / (self.world * bound)}}
        if probes.get("concurrent"):                       # the same bound from the rates measured under contention
            cR, cP = probes["concurrent"]["h2d_gbs"], probes["concurrent"]["pack_gbps"]
            f = pack_ratio
            cb = min(cR / max(1e-9, 1.0 - 0.75 * f), cP / f if f > 0 else float("inf"))
            cb = self.allmin(cb)
            e2e["roofline"]["peak_concurrent"] = self.world * cb
            e2e["roofline"]["frac_concurrent"] = value / (self.world * cb)
            e2e["roofline"]["concurrent_bound"] = ("at the packing ratio used: min(H2D rate / bytes per base over PCIe, packing rate / packed share), "
                                                   "both rates probed while H2D, D2H and the packers run together")
        # the same with the 2-bit transport switched off: every byte crosses PCIe as ASCII
        ctx.set_transport(0, 0.0)
        run()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e2e["plain_ascii_transport_value"] = W["total_bases"] * 2 / self.allmax(time.perf_counter() - t0) / 1e9
        ctx.set_transport(pack_threads, pack_ratio)
        # the caller already holds 2-bit packed reads (s2k_run_packed2) -- no host packing, a quarter of
        # the PCIe bytes.  Packed once outside the timed region; pinned like the ASCII buffer.
        if not W["split_one"]:
            hp = torch.from_numpy(ctx.pack2(hb_np, 16)).pin_memory()
            hp_np = hp.numpy()
            runp = lambda: ctx.run(hp_np, hso_np, L_PARAM, k, d, S.HashMode(mode), S.HashVariant(var), copy=False, packed2=True)
            runp()
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                outp = runp()
            torch.cuda.synchronize()
            dt3 = self.allmax(time.perf_counter() - t0)
            assert outp.n_items == n_items
            e2e["packed2_input_value"] = W["total_bases"] * 3 / dt3 / 1e9
            e2e["packed2_h2d_bytes_per_step"] = int(ctx.last_transport()[0])
            del hp
        return e2e

    def free(self, W):
        W.pop("d_bases", None); W.pop("d_so", None)
        self.torch.cuda.empty_cache()


def long_sequence_parity(O, bases, h, s, e, rv, l, k, d, mode, var, threads, piece=6_000_000):
    """One long sequence: the oracle (single-threaded per sequence, whole item arrays in memory) cannot take 3.1 Gbp at
    once, so it runs on overlapping pieces, each cut on a homopolymer-run boundary and processed as a sequence of its own
    (results of l-mers that start at or after such a cut do not depend on anything before it).  Piece p covers
    [c_p, c_{p+1} + ov); the items whose first minimizer starts in [c_p, c_{p+1}) are compared with the oracle's, field
    by field.  Every item belongs to exactly one piece.  (ntHash2-31 has no tail rule; for ntHash1 simd profiles the
    pieces before the last cannot know it, so only lengths on which it does not fire are checked this way.)"""
    n = len(bases)
    hpc = mode in (1, 3)
    rate = 2.0 * d * (0.75 if hpc else 1.0)
    ov = int(64 * (k + 8) / max(rate, 1e-9) + 64 * l + 4096)
    cuts = [0]
    while cuts[-1] + piece < n:
        c = cuts[-1] + piece
        while hpc and c < n and bases[c] == bases[c - 1]:
            c += 1
        cuts.append(c)
    cuts.append(n)
    s64 = s.astype(np.int64)

    def check(p):
        lo, hi = cuts[p], cuts[p + 1]
        top = min(n, hi + ov)
        want = O.kminmers(bases[lo:top], l, k, d, mode, var)
        ws = want["start"].astype(np.int64) + lo
        last = p + 2 == len(cuts)
        a, b = np.searchsorted(s64, lo, "left"), (len(s64) if last else np.searchsorted(s64, hi, "left"))
        wa, wb = 0, (len(ws) if last else np.searchsorted(ws, hi, "left"))
        if b - a != wb - wa:
            return False, 0
        ok = (np.array_equal(h[a:b], want["hash"][wa:wb]) and np.array_equal(s64[a:b], ws[wa:wb]) and
              np.array_equal(e[a:b].astype(np.int64), want["end"][wa:wb].astype(np.int64) + lo) and np.array_equal(rv[a:b], want["rev"][wa:wb]))
        return bool(ok), int(b - a)

    with ThreadPoolExecutor(max_workers=max(1, min(threads, 16))) as ex:
        res = list(ex.map(check, range(len(cuts) - 1)))
    return all(r[0] for r in res), sum(r[1] for r in res), len(cuts) - 1


def run_b200(args):
    from oracle import oracle as O   # parity checker and cpu_baseline only (never on the measured path)
    B = Bench(args)
    torch = B.torch
    rank, world = B.rank, B.world
    if args.workload == "c5":
        return run_c5(B, O, args)
    main_scaling = "strong" if (args.workload in ("c4", "c4off") and world > 1) else "weak"
    W = B.setup(args.workload, main_scaling, args.reads, args.mode)
    config = workload_config(args.workload, world, args.reads, args.mode)
    if W["split_one"]:
        config["sharding"] = f"one sequence cut into {world} base ranges, overlap-and-trim by ownership, no exchange of bases"
    R = B.time_resident(W, args.steps, args.warmup, sample_clocks=True)
    res, n_items = R["res"], R["n_items_rank"]
    roofline = R["roofline"]
    if args.workload == "c2" and W["mode"] == 3 and W["variant"] == 0:
        # DRAM traffic of the kernel from the committed `ncu --set full` capture of this build (1.000 Gbp launch of
        # this workload shape), scaled to this launch; instruction-issue context from the same capture.
        roofline["traffic"] = (NCU_DRAM_READ + NCU_DRAM_WRITE) * (W["n_bases"] / 1e9)
        roofline["traffic_source"] = f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, {NCU_SOURCE}, scaled by bases"
        clocks = R["clocks"]
        mhz = float(clocks.get("sm_mhz") or 1965.0) if isinstance(clocks, dict) else 1965.0
        sms = torch.cuda.get_device_properties(B.dev).multi_processor_count
        issue_peak = sms * 4 * 32 * mhz * 1e6            # thread-instructions per second
        rate = W["n_bases"] / (roofline["ms_per_step_in_kernel"] * 1e-3)
        roofline["instruction_bound"] = {"thread_instr_per_base": NCU_INSTR_PER_BASE, "source": NCU_SOURCE,
                                         "issue_frac": rate * NCU_INSTR_PER_BASE / issue_peak,
                                         "alu_pipe_frac": rate * NCU_INSTR_PER_BASE * NCU_ALU_SHARE / (issue_peak / 2)}

    # ---- host copy of the inputs (pinned): end-to-end runs and the parity check read it
    hb = torch.empty(W["n_bases"], dtype=torch.uint8).pin_memory()
    hb.copy_(W["d_bases"][:W["n_bases"]])
    hso = W["d_so"].cpu().pin_memory()
    hb_np, hso_np = hb.numpy(), hso.numpy().view(np.uint64)

    parity = None
    if not args.no_parity:
        res = B.step(W)                                    # a fresh result (e2e below reuses the context's buffers)
        parity = B.parity(O, W, res, hb_np)

    e2e = None
    if not args.no_e2e:
        probes = B.host_probes(hb_np, hb)
        pack_threads, pack_ratio, bound = B.choose_transport(probes)
        bound = B.allmin(bound)                            # the slowest rank's feed bounds the job
        e2e = B.e2e(W, hb_np, hso_np, n_items, probes, pack_threads, pack_ratio, bound)
    del hb, hso, hb_np, hso_np
    B.free(W)

    # ---- the other single-GPU BASELINE shapes in the same run, split over the ranks (strong scaling)
    extra = {}
    if args.workload == "c2" and not args.no_extra and not args.reads:
        for name in ("c3", "c4"):
            X = B.setup(name, "strong")
            RX = B.time_resident(X, max(2, min(args.steps, 5)), 3)
            px = None
            if not args.no_parity:
                px = B.parity(O, X, B.step(X))
            extra[name] = {"workload": X["desc"], "value": RX["value"], "unit": "Gbp/s", "ms_per_step": RX["ms_per_step"],
                           "scaling": "strong" if world > 1 else "single GPU", "n_gpus": world, "items_per_step": RX["items"],
                           "roofline": {k2: RX["roofline"][k2] for k2 in ("frac", "achieved", "peak", "ms_per_step_in_kernel", "window_stage_ms", "bytes_per_base")},
                           "gpu_launches": RX["launches"], "parity": px}
            B.free(X)

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        g, dt, sample = cpu_port_rate(O, W["L"], W["seed"], W["mode"], W["variant"], B.threads, target_s=12.0)
        cpu = {"value": g, "unit": "Gbp/s", "cores": B.threads, "kind": "port",
               "sample": sample + "; one iterator per read on all host threads (src/main.rs:65-79)"}

    if rank == 0:
        line = {"metric": "input Gbp/s -> k-min-mers", "value": R["value"], "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": R["ms_per_step"], "higher_is_better": True,
                "scaling": main_scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
                "minimizer_stream": ("not materialised: window stage reads the records in place (S2K_NO_MINIMIZER_STREAM)"
                                     if W["in_place"] else "ordered copy materialised (result.minimizers)"),
                "clocks": R["clocks"], "e2e": e2e, "gpu_launches": R["launches"], "roofline": roofline, "parity": parity,
                "cpu_baseline": cpu, "items_per_step": R["items"], "minimizers_per_step": R["minimizers"], "extra": extra or None}
        print(json.dumps(line))
    if world > 1:
        B.dist.destroy_process_group()
    B.ctx.close()
    return 0 if (parity is None or parity.get("digest_match", True) or not parity.get("checked")) else 1


# Figures of the committed ncu capture of the current k_minimizers (profiles/r2_last_kernel_summary.txt, 1.000 Gbp launch)
NCU_SOURCE = "profiles/r2_last_kernel_summary.txt"
NCU_DRAM_READ, NCU_DRAM_WRITE = 1.1013e9, 0.3088e9
NCU_INSTR_PER_BASE, NCU_ALU_SHARE = 29.87, 0.534      # smsp__inst_executed; ALU-class share (profiles/r2_last_phases.txt)


def run_c5(B, O, args):
    """BASELINE configs[4]: 30 Gbp of synthetic HiFi reads (1.5 M x 20 kb) as 3 slabs of 10 Gbp per GPU-pass, swept over
    density {0.001, 0.002, 0.005, 0.01} x k {5..10} x HPC {on, off}; strong scaling (the 3 slabs x reads are split over
    the ranks).  CPU-port column: once per (density, HPC) -- the CPU cost does not depend on k -- on a bounded sample.
    Parity: digest of every read of the first slab at the four corners k in {5, 10} x d in {0.001, 0.01} per HPC setting."""
    torch = B.torch
    rank, world = B.rank, B.world
    L, seed = 20000, 0x5EED0005
    total_reads = args.reads or 1_500_000
    slabs = 3
    reads_slab = total_reads // slabs // world             # per rank per slab
    WORKLOADS["c5"] = (L, reads_slab, seed, 3, 0, "configs[4]: 30 Gbp synthetic HiFi sweep")
    rows = []
    for hpc in (True, False):
        mode = 3 if hpc else 2
        cpu_by_d = {}
        for density in (0.001, 0.002, 0.005, 0.01):
            if rank == 0 and not args.no_cpu:
                g, _, _ = cpu_port_rate(O, L, seed, mode, 0, B.threads, target_s=3.0, density=density)
                cpu_by_d[density] = g
            for k in range(5, 11):
                ms_tot, items_tot, par = 0.0, 0, None
                for sl in range(slabs):
                    X = B.setup("c5", "weak", reads_slab, mode, k=k, density=density)
                    # the slab's own slice of the stream: (slab, rank) -> first base
                    X["first_base"] = (sl * world + rank) * reads_slab * L
                    B.ctx.synth_device(seed, X["first_base"], X["n_bases"], X["d_bases"].data_ptr())
                    torch.cuda.synchronize()
                    RX = B.time_resident(X, 2, 3)
                    ms_tot += RX["ms_per_step"]
                    items_tot += RX["items"]
                    if sl == 0 and not args.no_parity and k in (5, 10) and density in (0.001, 0.01):
                        par = B.parity(O, X, B.step(X))
                    B.free(X)
                n_bases = slabs * world * reads_slab * L
                row = {"hpc": hpc, "density": density, "k": k, "n_gpus": world, "Gbp": n_bases / 1e9, "ms": ms_tot,
                       "Gbp_per_s": n_bases / ms_tot / 1e6, "items": items_tot,
                       "bytes_per_base": (n_bases + 8 * (slabs * world * reads_slab + 1) + 17 * items_tot) / n_bases,
                       "cpu_port_Gbp_per_s": cpu_by_d.get(density), "parity": par}
                rows.append(row)
                if rank == 0:
                    print(json.dumps(row), flush=True)
    if rank == 0:
        print("| HPC | density | " + " | ".join(f"k={k}" for k in range(5, 11)) + " | CPU port (all host threads) |", file=sys.stderr)
        print("|---|---|" + "---|" * 7, file=sys.stderr)
        for hpc in (True, False):
            for density in (0.001, 0.002, 0.005, 0.01):
                r = [x for x in rows if x["hpc"] == hpc and x["density"] == density]
                cpu = r[0]["cpu_port_Gbp_per_s"]
                print(f"| {'on' if hpc else 'off'} | {density} | " + " | ".join(f"{x['Gbp_per_s']:.0f}" for x in r) +
                      (f" | {cpu:.1f} |" if cpu else " | - |"), file=sys.stderr)
        bad = [x for x in rows if x["parity"] and not x["parity"]["digest_match"]]
        print(f"parity corners checked: {sum(1 for x in rows if x['parity'])}, mismatches: {len(bad)}", file=sys.stderr)
    if world > 1:
        B.dist.destroy_process_group()
    B.ctx.close()
    return 0


def run_reference(args):
    """Reference arm: the CPU port of the reference's iterator on all host threads (rank 0 only)."""
    from oracle import oracle as O
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    O.build()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    name = args.workload if args.workload in WORKLOADS else "c2"
    L, n_reads, seed, mode, variant, desc = WORKLOADS[name]
    if args.mode >= 0:
        mode = args.mode
    config = workload_config(name, world, args.reads, args.mode)
    vals, sample = [], ""
    for _ in range(args.warmup):
        cpu_port_rate(O, L, seed, mode, variant, threads, target_s=1.0)
    for _ in range(args.steps):
        g, dt, sample = cpu_port_rate(O, L, seed, mode, variant, threads, target_s=8.0)
        vals.append((g, dt))
    v = float(np.mean([g for g, _ in vals]))
    line = {"impl": "reference", "metric": "input Gbp/s -> k-min-mers", "value": v, "unit": "Gbp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean([dt for _, dt in vals]) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": v, "unit": "Gbp/s", "cores": threads, "kind": "port",
                             "sample": "each step: " + sample + "; one iterator per read on all host threads (src/main.rs:65-79); "
                                       "the Rust crate itself is unbuildable here"},
            "e2e": {"value": v, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--reads", type=int, default=0, help="override the number of reads per GPU (debugging)")
    ap.add_argument("--mode", type=int, default=-1, help="override HashMode (0 Regular, 1 Hpc, 2 Simd, 3 HpcSimd)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-scale digest comparison with the CPU oracle")
    ap.add_argument("--no-extra", action="store_true", help="skip the c3 / c4 shapes after the main workload")
    ap.add_argument("--minimizer-stream", choices=["in-place", "ordered"], default="in-place",
                    help="ordered: also materialise the ordered minimizer stream (result.minimizers); in-place "
                         "(S2K_NO_MINIMIZER_STREAM): the window stage reads the records where the minimizer kernel left them")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
