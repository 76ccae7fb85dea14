#!/usr/bin/env python
"""bench.py -- input Gbp/s -> k-min-mers on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4]

A step = one pass of the whole hot path (HPC -> ntHash -> threshold -> ordered minimizers -> k-window hash)
over one batch of synthetic reads.  N=1 workload = BASELINE.json configs[1]: 500 000 HiFi-like reads x 20 kb
(10 Gbp), HPC on, ntHash1, l=31 k=5 density=0.01, HashMode::HpcSimd.  Multi-GPU: reads are sharded by rank,
no collective on the data path (weak scaling: each rank processes a full per-GPU workload).

  value    : device-resident throughput, inputs already in HBM, CUDA events, max over ranks
  e2e      : same metric through s2k_run with pinned HOST buffers (H2D of bases+offsets and D2H of all items
             inside the timed region)
  roofline : HBM roofline of the dominant kernel (k_minimizers): algorithmic bytes / its CUDA-event time
  cpu_baseline / --impl reference : the CPU oracle port of the reference's iterator on all host cores
             (the Rust reference itself cannot be built in this image: no cargo, nightly-only, git deps)
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
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (read_len, n_reads, seed, mode, variant, description)
    "c2": (20000, 500_000, 0x5EED0002, 3, 0, "configs[1]: synthetic HiFi-like reads 20 kb x 500k (10 Gbp), HPC on, ntHash1, HashMode::HpcSimd"),
    "c3": (150, 100_000_000, 0x5EED0003, 3, 0, "configs[2]: synthetic 150 bp reads x 100M (15 Gbp), HPC on, ntHash1, HashMode::HpcSimd"),
    "c4": (3_100_000_000, 1, 0x5EED0004, 3, 1, "configs[3]: one 3.1 Gbp sequence, chunked with halos, ntHash2-31, HPC on"),
}
L_PARAM, K_PARAM, DENSITY = 31, 5, 0.01


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


def cpu_port_rate(O, L, seed, mode, variant, threads, target_s=12.0, first_read=0):
    """Times the oracle port (mirrors src/main.rs:65-79: one iterator per read, count items) on a bounded sample."""
    probe_reads = max(threads * 4, int(2_000_000 // max(L, 1)) + 1) if L < 10_000_000 else 1
    probe_len = L if L < 10_000_000 else 20_000_000

    use_avx = variant == 0 and mode in (2, 3) and O.has_avx512()

    def run(n_reads, read_len):
        bases = O.synth(seed, first_read * L, n_reads * read_len)
        so = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
        t0 = time.perf_counter()
        if use_avx:      # AVX-512 restatement of the reference's vector path (oracle/s2k_cpu_avx512.c)
            r = O.avx512_batch(bases, so, L_PARAM, K_PARAM, DENSITY, mode, threads=threads, want_counts=False)
        else:            # scalar restatement (oracle/s2k_oracle.c)
            r = O.batch(bases, so, L_PARAM, K_PARAM, DENSITY, mode, variant, threads=threads, want_counts=False)
        return time.perf_counter() - t0, n_reads * read_len, r["total"]

    # Bounded sample: one slab of at most ~2 Gbp, repeated until about target_s seconds of CPU work have been timed.
    dt, nb, _ = run(probe_reads, probe_len)
    rate = nb / dt
    n_reads = max(probe_reads, min(int(rate * target_s / probe_len), int(2e9 // probe_len) or 1))
    bases = O.synth(seed, first_read * L, n_reads * probe_len)
    so = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(probe_len)
    call = (lambda: O.avx512_batch(bases, so, L_PARAM, K_PARAM, DENSITY, mode, threads=threads, want_counts=False)) if use_avx \
        else (lambda: O.batch(bases, so, L_PARAM, K_PARAM, DENSITY, mode, variant, threads=threads, want_counts=False))
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


IN_PLACE_DEFAULT = "in-place"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override the number of reads per GPU (debugging)")
    ap.add_argument("--mode", type=int, default=-1, help="override HashMode (0 Regular, 1 Hpc, 2 Simd, 3 HpcSimd)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--minimizer-stream", choices=["in-place", "ordered"], default=IN_PLACE_DEFAULT,
                    help="ordered: also materialise the ordered minimizer stream (result.minimizers); in-place "
                         "(S2K_NO_MINIMIZER_STREAM): the window stage reads the records where the minimizer kernel left them")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    L, n_reads, seed, mode, variant, desc = WORKLOADS[args.workload]
    if args.reads:
        n_reads = args.reads
    if args.mode >= 0:
        mode = args.mode
    threads = os.cpu_count() or 1
    config = {"workload": desc, "read_len": L, "reads_per_gpu": n_reads, "l": L_PARAM, "k": K_PARAM, "density": DENSITY,
              "hash_mode": ["Regular", "Hpc", "Simd", "HpcSimd"][mode], "hash": "ntHash2-31" if variant else "ntHash1-32",
              "sharding": f"reads split by rank x{world}, no data-path collective",
              "l2": "inputs (>=3 GB per step) far exceed the 126 MB L2; no flush needed"}

    from oracle import oracle as O   # cpu_baseline / reference arm only (checker, never the measured product path)

    # ------------------------------------------------------------------ reference arm: CPU port on host cores
    if args.impl == "reference":
        if rank != 0:
            return 0
        O.build()
        vals = []
        sample = ""
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

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    S = importlib.import_module("rust-seq2kminmers_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = S.Context(local_rank)
    ctx.set_timing(True)
    # Host transport per rank (s2k_ctx_set_transport): packing pays while a GPU's own PCIe link is the limit; with
    # several ranks on one host the host's memory bandwidth is, and packing only adds traffic.
    pack_threads, pack_ratio = (0, 0.7) if world == 1 else ((6, 0.5) if world == 2 else (0, 0.0))
    ctx.set_transport(pack_threads, pack_ratio)
    split_one = args.workload == "c4" and world > 1    # one sequence cut into base ranges: strong scaling
    sharding = importlib.import_module("rust-seq2kminmers_b200.sharding")
    if split_one:
        b0, b1, lo, hi = sharding.sequence_ranges(L, world, sharding.default_overlap_right(L_PARAM, K_PARAM, DENSITY))[rank]
        n_bases, first_base = hi - lo, lo
        config["sharding"] = f"one sequence cut into {world} base ranges, overlap-and-trim by ownership, no exchange of bases"
    else:
        n_bases, first_base = L * n_reads, rank * n_reads * L   # each rank owns its own slice of the synthetic stream
    d_bases = torch.empty(n_bases + 16, dtype=torch.uint8, device=dev)
    ctx.synth_device(seed, first_base, n_bases, d_bases.data_ptr())
    if split_one:
        d_so = torch.tensor([0, n_bases], dtype=torch.int64, device=dev)
    else:
        d_so = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * L
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream

    in_place = args.minimizer_stream == "in-place" and not split_one      # a rank of a split sequence needs the stream
    config["minimizer_stream"] = ("not materialised: window stage reads the records in place (S2K_NO_MINIMIZER_STREAM)"
                                  if in_place else "ordered copy materialised (result.minimizers)")

    def step():
        return ctx.run_device(d_bases.data_ptr(), d_so.data_ptr(), n_reads, n_bases, L_PARAM, K_PARAM, DENSITY,
                              S.HashMode(mode), S.HashVariant(variant), stream=stream, no_tail_rule=split_one,
                              no_minimizer_stream=in_place)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        res = step()
    n_items, n_min = int(res.n_items), int(res.n_minimizers)
    if split_one:                                      # what this rank OWNS: windows whose first minimizer starts in [b0, b1)
        mins_t = torch.as_tensor(S.DeviceArray(res.minimizers, n_min * 16, "|u1"), device=dev).view(torch.int32).view(-1, 4)
        starts = (mins_t[:, 1].to(torch.int64) & 0xffffffff) + lo
        i0 = int(torch.searchsorted(starts, torch.tensor([b0], device=dev))[0])
        i1 = n_min if rank == world - 1 else int(torch.searchsorted(starts, torch.tensor([b1], device=dev))[0])
        n_items, n_min = max(0, min(i1, max(0, n_min - K_PARAM + 1)) - i0), i1 - i0
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kms, wms, kl = [], [], 0
    ev[0].record()
    for _ in range(args.steps):
        res = step()
        a, b, c = ctx.last_kernel_ms()
        kms.append(a); wms.append(b); kl = c
    ev[1].record()
    barrier()
    launches = ctx.launch_count - launches0
    ms_total = ev[0].elapsed_time(ev[1])
    clocks = sampler.summary()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = (L if split_one else world * n_bases) / (ms_step * 1e-3) / 1e9

    # optional final gather of per-GPU counts over NCCL (the only collective; not on the data path)
    if world > 1:
        per_rank, _first_item = sharding.gather_totals(n_items, n_min, device=dev)
        counts = torch.tensor(per_rank.sum(axis=0))
    else:
        counts = torch.tensor([n_items, n_min], dtype=torch.int64)

    # roofline of the dominant kernel (k_minimizers): algorithmic bytes = bases + offsets + 17 B per item
    alg_bytes = n_bases + 8 * (n_reads + 1) + 17 * n_items
    k_ms = float(np.mean(kms))
    peak, peak_src = peaks()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    # DRAM traffic of the kernel from the committed `ncu --set full` capture (profiles/r1_final_kernel_summary.txt:
    # 1.1076 GB read + 0.3191 GB written for a 1.000 Gbp launch of this workload shape), scaled to this launch.
    traffic = (1.1076e9 + 0.3191e9) * (n_bases / 1e9) if args.workload == "c2" and mode == 3 and variant == 0 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r1_final_kernel_summary.txt, scaled by bases" if traffic else None,
                "algorithmic_bytes": alg_bytes,
                "kernel": "k_minimizers", "launches_per_step": kl, "ms_per_step_in_kernel": k_ms,
                "window_stage_ms": float(np.mean(wms)), "bytes_per_base": alg_bytes / n_bases, "peak_source": peak_src}
    if traffic:
        # What actually bounds the kernel (context for the HBM fraction above): instruction
        # throughput.  35.3 thread-instructions per base, 52 % of them on the ALU pipe (16 lanes per SM sub-partition),
        # from the committed ncu capture; ceilings from the SM count and the sampled clock.
        mhz = float(clocks.get("sm_mhz") or 1965.0) if isinstance(clocks, dict) else 1965.0
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = sms * 4 * 32 * mhz * 1e6            # thread-instructions per second
        ipb, alu_share = 35.1, 0.52
        rate = n_bases / (k_ms * 1e-3)
        roofline["instruction_bound"] = {"thread_instr_per_base": ipb, "source": "profiles/r1_final_kernel_summary.txt",
                                         "issue_frac": rate * ipb / issue_peak, "alu_pipe_frac": rate * ipb * alu_share / (issue_peak / 2)}

    # ------------------------------------------------------------------ end to end through s2k_run (host buffers)
    e2e = None
    if not args.no_e2e:
        hb = torch.empty(n_bases, dtype=torch.uint8).pin_memory()
        hb.copy_(d_bases[:n_bases])
        hso = d_so.cpu().pin_memory()
        hb_np, hso_np = hb.numpy(), hso.numpy().view(np.uint64)
        e_steps = args.steps
        out = ctx.run(hb_np, hso_np, L_PARAM, K_PARAM, DENSITY, S.HashMode(mode), S.HashVariant(variant), copy=False,
                      no_tail_rule=split_one)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            out = ctx.run(hb_np, hso_np, L_PARAM, K_PARAM, DENSITY, S.HashMode(mode), S.HashVariant(variant), copy=False,
                          no_tail_rule=split_one)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        assert split_one or out.n_items == n_items
        h2d_actual, n_packed, n_plain = ctx.last_transport()
        e2e = {"value": (L if split_one else world * n_bases) * e_steps / float(tt.item()) / 1e9, "unit": "Gbp/s",
               "h2d_bytes_per_step": int(h2d_actual),
               "host_input_bytes_per_step": n_bases + 8 * (n_reads + 1),
               "d2h_bytes_per_step": 17 * int(out.n_items) + 8 * (n_reads + 1) * 2 + 4 * n_reads,
               "steps": e_steps, "api": "s2k_run (C ABI, pinned host ASCII buffers in, pinned host items out)",
               "transport": f"{n_packed} slabs packed to 2 bits/base by host threads + {n_plain} slabs as plain ASCII (s2k_ctx_set_transport: threads {pack_threads or 'default'}, ratio {pack_ratio})"}
        # the same with the 2-bit transport switched off: every byte crosses PCIe as ASCII
        ctx.set_transport(0, 0.0)
        ctx.run(hb_np, hso_np, L_PARAM, K_PARAM, DENSITY, S.HashMode(mode), S.HashVariant(variant), copy=False, no_tail_rule=split_one)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            ctx.run(hb_np, hso_np, L_PARAM, K_PARAM, DENSITY, S.HashMode(mode), S.HashVariant(variant), copy=False, no_tail_rule=split_one)
        torch.cuda.synchronize()
        tt2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt2, op=dist.ReduceOp.MAX)
        e2e["plain_ascii_transport_value"] = (L if split_one else world * n_bases) * 2 / float(tt2.item()) / 1e9
        ctx.set_transport(pack_threads, pack_ratio)
        # the caller already holds 2-bit packed reads (s2k_run_packed2) -- no host packing, a quarter of
        # the PCIe bytes.  Packed once outside the timed region; pinned like the ASCII buffer.
        if not split_one:
            hp = torch.from_numpy(ctx.pack2(hb_np, 16)).pin_memory()
            hp_np = hp.numpy()
            ctx.run(hp_np, hso_np, L_PARAM, K_PARAM, DENSITY, S.HashMode(mode), S.HashVariant(variant), copy=False, packed2=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                outp = ctx.run(hp_np, hso_np, L_PARAM, K_PARAM, DENSITY, S.HashMode(mode), S.HashVariant(variant), copy=False, packed2=True)
            torch.cuda.synchronize()
            tt3 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt3, op=dist.ReduceOp.MAX)
            assert outp.n_items == n_items
            e2e["packed2_input_value"] = world * n_bases * 3 / float(tt3.item()) / 1e9
            e2e["packed2_h2d_bytes_per_step"] = int(ctx.last_transport()[0])
            del hp
        del hb, hso

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        g, dt, sample = cpu_port_rate(O, L, seed, mode, variant, threads, target_s=12.0)
        cpu = {"value": g, "unit": "Gbp/s", "cores": threads, "kind": "port",
               "sample": sample + "; one iterator per read on all host threads (src/main.rs:65-79)"}

    if rank == 0:
        line = {"metric": "input Gbp/s -> k-min-mers", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if split_one else "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "items_per_step": int(counts[0].item()), "minimizers_per_step": int(counts[1].item())}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
