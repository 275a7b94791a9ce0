#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "$@"; do
  for cfg in ${AB_CFGS:-c2plane c4}; do
  env $v timeout 600 python bench.py --config $cfg --no-cpu --no-extra --no-e2e --reps 5 2> gpurun_out/abenv.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ex=d['extra']
print('[$v] $cfg value %.1f  b2b %.4f ms  first %.3f last %.3f  stages %s' % (d['value'], ex['back_to_back_ms_per_step'], ex['first_iteration_ms'], ex['last_iteration_ms'], {k: round(x,4) for k,x in ex['stage_ms_per_iteration'].items() if x}))"
  done
done
