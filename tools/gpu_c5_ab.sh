#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "$@"; do
  env $v timeout 300 python bench.py --config c5 --no-cpu 2> gpurun_out/c5ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$v] c5 %.1f pairs/s  %.1f it/s  launches %d' % (d['extra']['pairs_per_s'], d['value'], d['gpu_launches']))"
done
tail -2 gpurun_out/c5ab.err
