#!/bin/bash
# Final source at 2 GPUs, launched as the driver launches it (both arms), plus the 2-rank counting path.
mkdir -p gpurun_out
run() { N=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N "$@" 2>>gpurun_out/r2_last_2gpu.err | tail -1; }
run 2 > gpurun_out/r2_last_bench_2gpu.json
run 2 --impl reference > gpurun_out/r2_last_ref_2gpu.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_last_bench_2gpu.json")); e=d.get("e2e") or {}
print("2 GPUs value", round(d["value"],1), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", e.get("value"), "parity", (d.get("parity") or {}).get("digest_match"), "extra", {k: round(v["value"],1) for k,v in (d.get("extra") or {}).items()})
r=json.load(open("gpurun_out/r2_last_ref_2gpu.json")); print("ref arm", r.get("value"), r.get("n_gpus"))
PY
tail -3 gpurun_out/r2_last_2gpu.err
