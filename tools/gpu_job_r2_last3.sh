#!/bin/bash
# Config-5 sweep (30 Gbp per point) on the final source, with the CPU-port column and corner parity.
mkdir -p gpurun_out
timeout 700 python bench.py --workload c5 --no-e2e 2>gpurun_out/r2_last_c5.err > gpurun_out/r2_last_config5_sweep.jsonl; echo "c5 rc=$?"
tail -12 gpurun_out/r2_last_c5.err
python tools/c5_table.py gpurun_out/r2_last_config5_sweep.jsonl > gpurun_out/r2_last_config5_sweep.md
