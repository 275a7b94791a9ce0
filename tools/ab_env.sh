#!/bin/bash
# A/B over an environment variable: tools/ab_env.sh VAR v1 v2 ...
VAR=$1; shift
for V in "$@"; do
  export $VAR=$V
  timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/bench_env.json 2> gpurun_out/bench_env.err || tail -3 gpurun_out/bench_env.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_env.json"))
pp = d["extra"]["point_to_plane"]
print("$VAR=%-4s p2point it/s %7.1f (first/last iter ms %.3f/%.3f) | p2plane it/s %7.1f" % ("$V", d["value"], d["extra"]["per_iteration_ms_first_last"][0], d["extra"]["per_iteration_ms_first_last"][1], pp["iterations_per_s"]))
PY
done
