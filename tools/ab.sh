#!/bin/bash
# A/B harness for kernel experiments on the GPU box: builds the library once per variant (extra nvcc flags) and prints
# the resident throughput and the minimizer-kernel time of the default bench.  Usage: tools/ab.sh "" "-DS2K_X=1" ...
for v in "$@"; do
  S2K_NVCC_EXTRA="$v" python -c "import __graft_entry__ as g; g.build_cuda(True)" > /dev/null 2>&1 || { echo "[$v] build failed"; continue; }
  for rep in 1 2; do
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$v]', round(d['value'],1), 'Gbp/s  k_minimizers', round(d['roofline']['ms_per_step_in_kernel'],3), 'ms  items', d['items_per_step'])"
  done
done 2>&1 | tee -a gpurun_out/ab.txt
