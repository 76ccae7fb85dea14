#!/bin/bash
# Config 3 (150-bp reads): A/B of the window stage's same-sequence test, then one ncu --set full capture of k_minimizers on
# 10^7 reads with the per-stage instruction table (where do the extra 17 % per base go on short reads?).
mkdir -p gpurun_out
line() { python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity --no-extra "${@:2}" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$1]', round(d['value'],1), 'Gbp/s  step', round(d['ms_per_step'],3), 'k_min', round(d['roofline']['ms_per_step_in_kernel'],3), 'win', round(d['roofline']['window_stage_ms'],4), 'items', d['items_per_step'])" | tee -a gpurun_out/ab_c3.txt; }
line "ridtest c3" --workload c3
line "ridtest c2"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_minimizers --launch-skip 3 -c 1 -f -o gpurun_out/r2_c3_k1 \
  python bench.py --workload c3 --reads 10000000 --steps 1 --warmup 3 --no-cpu --no-e2e --no-parity --no-extra > gpurun_out/r2_c3_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_phases.py gpurun_out/r2_c3_k1.ncu-rep 1.5e9 > gpurun_out/r2_c3_phases.txt 2>&1
python tools/ncu_lines.py gpurun_out/r2_c3_k1.ncu-rep 1.5e9 40 > gpurun_out/r2_c3_kernel_summary.txt 2>&1
S2K_NVCC_EXTRA="-DS2K_WIN_NO_RIDTEST" python -c "import __graft_entry__ as g; g.build_cuda(True)" > /dev/null 2>&1 || echo "build failed"
line "no ridtest c3" --workload c3
line "no ridtest c2"
