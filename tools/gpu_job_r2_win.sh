#!/bin/bash
# A/B of the in-place window stage (k_windows_t): registers per thread (launch bounds) x blocks per SM in the grid.
# (History: S2K_WIN_BPS was a temporary environment knob of run_device for this A/B; the grid is fixed at 16 CTAs per SM now.)
mkdir -p gpurun_out
line() { python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$1]', round(d['value'],1), 'Gbp/s  step', round(d['ms_per_step'],3), 'k_min', round(d['roofline']['ms_per_step_in_kernel'],3), 'win', round(d['roofline']['window_stage_ms'],4))" | tee -a gpurun_out/ab_win.txt; }
for bps in 16 6; do S2K_WIN_BPS=$bps line "win default bps=$bps"; done
S2K_NVCC_EXTRA="-DS2K_WIN_MINB=6" python -c "import __graft_entry__ as g; g.build_cuda(True)" > /dev/null 2>&1 || echo "build failed"
for bps in 16 6; do S2K_WIN_BPS=$bps line "win MINB=6 bps=$bps"; done
S2K_NVCC_EXTRA="-DS2K_WIN_MINB=8" python -c "import __graft_entry__ as g; g.build_cuda(True)" > /dev/null 2>&1 || echo "build failed"
for bps in 16 8; do S2K_WIN_BPS=$bps line "win MINB=8 bps=$bps"; done
timeout 300 python -m pytest tests -m gpu -x -q -k "h16" 2>&1 | tail -2
python tools/bench_h64.py 100000 2>&1 | tee gpurun_out/r2_h16.txt
