#!/bin/bash
# ncu --set full of the in-place window stage on the final source (after the L2 prefetch), config 2.
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_windows_t --launch-skip 3 -c 1 -f -o gpurun_out/r2_last_win \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-parity --no-extra > gpurun_out/r2_last_win_ncu.log 2>&1; echo "ncu rc=$?"
