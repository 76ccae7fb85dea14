mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log)
tail -5 gpurun_out/gpu_tests.log
for ms in ordered in-place ordered in-place; do
  timeout 120 python bench.py --no-cpu --no-e2e --minimizer-stream $ms 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$ms', round(d['value'],1), 'Gbp/s  step', round(d['ms_per_step'],3), 'ms  k_min', round(d['roofline']['ms_per_step_in_kernel'],3), 'win', round(d['roofline']['window_stage_ms'],3), 'items', d['items_per_step'])" | tee -a gpurun_out/ab_in_place.txt
done
for w in c3 c4; do for ms in ordered in-place; do
  timeout 200 python bench.py --no-cpu --no-e2e --workload $w --minimizer-stream $ms 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w $ms', round(d['value'],1), 'Gbp/s  step', round(d['ms_per_step'],3), 'ms  items', d['items_per_step'])" | tee -a gpurun_out/ab_in_place.txt
done; done
timeout 300 python tools/sweep_c5.py > gpurun_out/sweep_c5.jsonl 2> gpurun_out/sweep_c5.md; echo sweep rc=$?; tail -10 gpurun_out/sweep_c5.md
