#!/bin/bash
# Config 3: one loop over a thread's sequence starts (all four pieces) against the loop per piece; parity tests on the new build.
mkdir -p gpurun_out
line() { python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra "${@:2}" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$1]', round(d['value'],1), 'Gbp/s  step', round(d['ms_per_step'],3), 'k_min', round(d['roofline']['ms_per_step_in_kernel'],3), 'win', round(d['roofline']['window_stage_ms'],4), 'items', d['items_per_step'], 'parity', (d.get('parity') or {}).get('digest_match'))" | tee -a gpurun_out/ab_c3.txt; }
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
line "merged start loop c3" --workload c3
line "merged start loop c2"
line "merged start loop c5-like: c2 mode Simd" --mode 2 --no-parity
S2K_NVCC_EXTRA="-DS2K_START_LOOP_PER_PIECE" python -c "import __graft_entry__ as g; g.build_cuda(True)" > /dev/null 2>&1 || echo "build failed"
line "loop per piece c3" --workload c3 --no-parity
line "loop per piece c2" --no-parity
line "loop per piece: c2 mode Simd" --mode 2 --no-parity
