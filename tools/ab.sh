#!/bin/bash
# A/B of tuning variants: bench each libpmgpu_<variant>.so, print the stage times
for V in "$@"; do
  if [ "$V" = "base" ]; then unset PMGPU_VARIANT; else export PMGPU_VARIANT=$V; fi
  timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err || tail -3 gpurun_out/bench_$V.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_$V.json"))
pp = d["extra"]["point_to_plane"]
print("%-8s p2point it/s %7.1f knn %.3f ms | p2plane it/s %7.1f knn %.3f ms" % ("$V", d["value"], d["extra"]["stage_ms_per_iteration"]["knn"], pp["iterations_per_s"], pp["stage_ms_per_iteration"]["knn"]))
PY
done
