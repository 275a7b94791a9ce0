#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for v in "A=1" "PMGPU_NO_PDL=1"; do
  env $v timeout 600 python bench.py --no-cpu --no-extra --reps 5 --e2e-reps 9 2> gpurun_out/pdl.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ex=d['extra']
print('[$v] value %.1f  b2b %.4f ms  stages %s  e2e %.1f (%.3f ms) py %s' % (d['value'], ex['back_to_back_ms_per_step'], {k: round(x,4) for k,x in ex['stage_ms_per_iteration'].items() if x}, d['e2e']['value'], 1e3*d['e2e']['seconds_per_registration'], (ex.get('e2e_python_mirror') or ex.get('e2e_cpp_mirror') or {}).get('seconds_per_registration')))"
done
tail -2 gpurun_out/pdl.err
