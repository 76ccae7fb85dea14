mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_c3_launches.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --no-extra > gpurun_out/r2_c3_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/r2_c3_launches.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
agg=collections.OrderedDict()
data=rows[hdr+1:]
# last step = last 9ish launches: print the final 12 launches
for r in data[-14:]:
    print(r[4][:60], r[-1], r[-2] if len(r)>2 else "")
PY
for n in 19375 155000; do timeout 120 python bench.py --reads $n --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('reads $n', round(d['value'],1), 'Gbp/s step', round(d['ms_per_step'],4), 'ms k_min', round(d['roofline']['ms_per_step_in_kernel'],4), 'win', round(d['roofline']['window_stage_ms'],4), 'launches', d['gpu_launches'])"; done
