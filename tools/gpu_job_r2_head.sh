#!/bin/bash
# Validation of the head of round 2 on a fresh box, as the driver does it: parity tests, smoke(), the default bench line,
# the reference arm.  Output: gpurun_out/r2_head_*.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_head_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_head_gpu_tests.log)
tail -3 gpurun_out/r2_head_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py 2>gpurun_out/r2_head_bench.err | tail -1 > gpurun_out/r2_head_bench.json; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_head_bench.json"))
print("c2", round(d["value"],1), "Gbp/s step", round(d["ms_per_step"],3), "k_min", round(d["roofline"]["ms_per_step_in_kernel"],3), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],1), "parity", d.get("parity",{}).get("digest_match"))
for k,v in d.get("extra",{}).items(): print(k, round(v["value"],1), "frac", round(v["roofline"]["frac"],4), "parity", v.get("parity",{}).get("digest_match"))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
timeout 600 python bench.py --impl reference 2>gpurun_out/r2_head_ref.err | tail -1 > gpurun_out/r2_head_ref.json; echo "ref rc=$?"
cut -c1-300 gpurun_out/r2_head_ref.json
