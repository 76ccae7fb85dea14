#!/bin/bash
# one line per run of the default bench: resident Gbp/s, k_minimizers ms.  usage: [ENV=..] tools/bench_line.sh LABEL [bench args]
L=$1; shift
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$L]', round(d['value'],1), 'Gbp/s  k_minimizers', round(d['roofline']['ms_per_step_in_kernel'],3), 'ms  step', round(d['ms_per_step'],3), ' items', d['items_per_step'])" | tee -a gpurun_out/ab.txt
