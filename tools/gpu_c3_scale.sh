#!/bin/bash
# BASELINE configs[2] (10 M-point map, 1 M-point reading, point-to-plane, normals knn 20) with the queries sharded over N ranks and the
# map's normals computed per slice + all-gathered (gpurun --gpus N)
set -u
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2978$N bench.py --config c3 --gpus $N --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2e_bench_c3_n$N.json 2> gpurun_out/r2e_bench_c3_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r2e_bench_c3_n$N.err | cut -c1-400
python - <<PY
import json
for line in open("gpurun_out/r2e_bench_c3_n$N.json"):
    if line.startswith("{"):
        d = json.loads(line); ex = d["extra"]
        print("c3 N=$N value", round(d["value"], 1), "e2e", d["e2e"] and (round(d["e2e"]["value"], 1), d["e2e"].get("seconds_per_registration")), "stages", ex["stage_ms_per_iteration"],
              "normals ms", ex.get("normals_ms_resident"))
PY
