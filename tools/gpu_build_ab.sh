#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for v in "A=1" "PMGPU_BUILD_SORT=1"; do
  env $v timeout 600 python bench.py --no-cpu --no-extra --reps 3 --e2e-reps 9 2> gpurun_out/bab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ex=d['extra']
print('[$v] value %.1f e2e %.1f (%.3f ms) other mirror %s' % (d['value'], d['e2e']['value'], 1e3*d['e2e']['seconds_per_registration'], (ex.get('e2e_python_mirror') or ex.get('e2e_cpp_mirror') or {}).get('seconds_per_registration')))"
  env $v timeout 600 python bench.py --config c3 --no-cpu --no-extra --reps 3 --e2e-reps 5 2>> gpurun_out/bab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('[$v] c3 value %.1f e2e %.1f (%.3f ms)' % (d['value'], d['e2e']['value'], 1e3*d['e2e']['seconds_per_registration']))"
done
tail -2 gpurun_out/bab.err
