#!/bin/bash
# Last GPU job of round 2 on the final source: parity tests, smoke, the default bench line exactly as the driver runs it
# (and the reference arm), the ncu launch list of the same command, one `ncu --set full` capture of k_minimizers (1 Gbp
# launch) with per-phase and per-line summaries.  Everything lands in gpurun_out/r2_last_*.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_last_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_last_gpu_tests.log)
tail -3 gpurun_out/r2_last_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py 2>gpurun_out/r2_last_bench.err | tail -1 > gpurun_out/r2_last_bench.json; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_last_bench.json"))
print("c2", round(d["value"],1), "Gbp/s step", round(d["ms_per_step"],3), "k_min", round(d["roofline"]["ms_per_step_in_kernel"],3), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],1), "parity", d.get("parity",{}).get("digest_match"))
for k,v in d.get("extra",{}).items(): print(k, round(v["value"],1), "frac", round(v["roofline"]["frac"],4), "parity", v.get("parity",{}).get("digest_match"))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
timeout 600 python bench.py --impl reference 2>gpurun_out/r2_last_ref.err | tail -1 > gpurun_out/r2_last_ref.json; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_last_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --no-extra > gpurun_out/r2_last_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_minimizers --launch-skip 3 -c 1 -f -o gpurun_out/r2_last_k1 \
  python bench.py --reads 50000 --steps 1 --warmup 3 --no-cpu --no-e2e --no-parity --no-extra > gpurun_out/r2_last_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_phases.py gpurun_out/r2_last_k1.ncu-rep 1e9 > gpurun_out/r2_last_phases.txt 2>&1
python tools/ncu_lines.py gpurun_out/r2_last_k1.ncu-rep 1e9 40 > gpurun_out/r2_last_kernel_summary.txt 2>&1
head -3 gpurun_out/r2_last_phases.txt
