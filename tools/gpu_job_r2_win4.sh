#!/bin/bash
# In-place window stage: also the NEXT tile's records prefetched into L2 (after the first pass of the current tile).
mkdir -p gpurun_out
line() { python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra "${@:2}" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$1]', round(d['value'],1), 'Gbp/s  step', round(d['ms_per_step'],3), 'k_min', round(d['roofline']['ms_per_step_in_kernel'],3), 'win', round(d['roofline']['window_stage_ms'],4), 'items', d['items_per_step'], 'parity', (d.get('parity') or {}).get('digest_match'))" | tee -a gpurun_out/ab_win4.txt; }
line "next tile prefetched too: c2"
line "next tile prefetched too: c3" --workload c3 --no-parity
S2K_NVCC_EXTRA="-DS2K_WIN_NO_NEXT" python -c "import __graft_entry__ as g; g.build_cuda(True)" > /dev/null 2>&1 || echo "build failed"
line "own tile only (previous commit): c2" --no-parity
line "own tile only (previous commit): c3" --workload c3 --no-parity
