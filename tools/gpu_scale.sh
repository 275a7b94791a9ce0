#!/bin/bash
# the default bench line at N ranks, as the driver launches it (gpurun --gpus N)
set -u
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2977$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r2e_bench_n$N.err | cut -c1-400
python - <<PY
import json
for line in open("gpurun_out/r2e_bench_n$N.json"):
    if line.startswith("{"):
        d = json.loads(line); ex = d["extra"]
        print("N=$N value", round(d["value"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), "stages", ex["stage_ms_per_iteration"])
        print("  c4", ex.get("c4_strong_scaling", {}).get("value"), ex.get("c4_strong_scaling", {}).get("stage_ms_per_iteration"))
        print("  pairs", ex.get("pairs_replicas", {}).get("value"), "pairs e2e", (ex.get("pairs_replicas", {}).get("e2e") or {}).get("value"))
PY
