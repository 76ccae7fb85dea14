mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_final.json
timeout 200 python bench.py --workload c3 --no-cpu > gpurun_out/bench_c3.json 2>/dev/null; echo "c3 rc=$?"
timeout 200 python bench.py --workload c4 --no-cpu > gpurun_out/bench_c4.json 2>/dev/null; echo "c4 rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
