#!/bin/bash
# Round-2 multi-GPU job (run under `gpurun --gpus 8`): the default bench line at 4 and 8 GPUs exactly as the driver
# launches it, config 3 at 8 GPUs and config 4 split over 8 GPUs (strong scaling).
mkdir -p gpurun_out
run() { N=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N "$@" 2>>gpurun_out/r2_scale.err | tail -1; }
for N in 4 8; do run $N --no-cpu > gpurun_out/r2_scale_c2_${N}gpu.json; done
run 8 --workload c3 --no-cpu --no-e2e --no-extra > gpurun_out/r2_scale_c3_8gpu.json
run 8 --workload c4 --no-cpu --no-e2e --no-extra > gpurun_out/r2_scale_c4_8gpu.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_scale_*.json")):
    try: d=json.load(open(f))
    except Exception as e: print(f, "unreadable", e); continue
    e=d.get("e2e") or {}
    print(f, "value", round(d["value"],1), d["unit"], "ms", round(d["ms_per_step"],3), "scaling", d["scaling"], "e2e", e.get("value"), "e2e frac", (e.get("roofline") or {}).get("frac"), "parity", (d.get("parity") or {}).get("digest_match"), "extra", {k: round(v["value"],1) for k,v in (d.get("extra") or {}).items()})
PY
tail -5 gpurun_out/r2_scale.err
