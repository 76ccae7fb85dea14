#!/bin/bash
# Final source (window stage with the L2 prefetch): parity tests, smoke, default bench line, launch list of the same command.
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_last_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_last_gpu_tests.log)
tail -3 gpurun_out/r2_last_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>gpurun_out/r2_last_bench.err | tail -1 > gpurun_out/r2_last_bench.json; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_last_bench.json"))
print("c2", round(d["value"],1), "Gbp/s step", round(d["ms_per_step"],3), "k_min", round(d["roofline"]["ms_per_step_in_kernel"],3), "win", round(d["roofline"]["window_stage_ms"],3), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],1), "parity", d.get("parity",{}).get("digest_match"))
for k,v in d.get("extra",{}).items(): print(k, round(v["value"],1), "frac", round(v["roofline"]["frac"],4), "parity", v.get("parity",{}).get("digest_match"))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_last_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --no-extra > gpurun_out/r2_last_ncu_list.log 2>&1; echo "ncu list rc=$?"
