#!/bin/bash
# In-place window stage: per-sequence prefixes cached across the passes of a tile (long reads) against a load round per pass.
mkdir -p gpurun_out
line() { python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra "${@:2}" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$1]', round(d['value'],1), 'Gbp/s  step', round(d['ms_per_step'],3), 'k_min', round(d['roofline']['ms_per_step_in_kernel'],3), 'win', round(d['roofline']['window_stage_ms'],4), 'items', d['items_per_step'], 'parity', (d.get('parity') or {}).get('digest_match'))" | tee -a gpurun_out/ab_win2.txt; }
line "ridcache, 5 CTAs/SM: c2"
line "ridcache, 5 CTAs/SM: c3" --workload c3
line "ridcache, 5 CTAs/SM: c4" --workload c4
S2K_NVCC_EXTRA="-DS2K_WIN_MINB=6" python -c "import __graft_entry__ as g; g.build_cuda(True)" > /dev/null 2>&1 || echo "build failed"
line "ridcache, 6 CTAs/SM (spills): c2" --no-parity
S2K_NVCC_EXTRA="-DS2K_WIN_MINB=6 -DS2K_WIN_NO_RIDCACHE" python -c "import __graft_entry__ as g; g.build_cuda(True)" > /dev/null 2>&1 || echo "build failed"
line "no ridcache, 6 CTAs/SM (previous commit): c2" --no-parity
line "no ridcache, 6 CTAs/SM (previous commit): c4" --workload c4 --no-parity
