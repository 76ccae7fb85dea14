#!/bin/bash
# Round-2 GPU job: parity tests, resident bench of config 2, one `ncu --set full` capture of k_minimizers (1 Gbp launch),
# per-phase and per-line summaries.  usage: tools/gpu_job_r2.sh TAG [skip-tests]
TAG=${1:-r2}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  (timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log)
  tail -4 gpurun_out/${TAG}_gpu_tests.log
fi
for rep in 1 2; do
timeout 200 python bench.py --no-cpu --no-e2e 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("${TAG} c2", round(d["value"],1), "Gbp/s step", round(d["ms_per_step"],3), "ms k_min", round(d["roofline"]["ms_per_step_in_kernel"],3), "win", round(d["roofline"]["window_stage_ms"],3), "frac", round(d["roofline"]["frac"],4), "items", d["items_per_step"])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_minimizers --launch-skip 3 -c 1 -f -o gpurun_out/${TAG}_k1 \
  python bench.py --reads 50000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu rc=$?"
python tools/ncu_phases.py gpurun_out/${TAG}_k1.ncu-rep 1e9 > gpurun_out/${TAG}_phases.txt 2>&1
python tools/ncu_lines.py gpurun_out/${TAG}_k1.ncu-rep 1e9 40 > gpurun_out/${TAG}_kernel_summary.txt 2>&1
head -20 gpurun_out/${TAG}_phases.txt
