#!/bin/bash
# second-stream reading upload: parity, then e2e A/B (c2plane and c5 pair stream)
set -u
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for v in "" "PMGPU_NO_OVERLAP=1"; do
  echo "== e2e c2plane $v"; env $v timeout 600 python bench.py --no-cpu --no-extra --reps 3 --e2e-reps 9 2> gpurun_out/ov.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/reg', round(1e3*d['e2e']['seconds_per_registration'],3))"
  echo "== c5 $v"; env $v timeout 600 python bench.py --config c5 --pairs 256 --no-cpu 2>> gpurun_out/ov.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 value', round(d['value'],1), 'pairs/s', round(d['extra']['pairs_per_s'],1))"
done
tail -3 gpurun_out/ov.err
