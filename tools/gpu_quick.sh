#!/bin/bash
# parity + stage times of the current build (+ the phase probe of the fused kernel when the prof variant is there)
set -u
mkdir -p gpurun_out
if [ -f libpointmatcher_b200/libpmgpu_prof.so ]; then echo "== probe"; PMGPU_VARIANT=prof timeout 200 python tools/probe_fused_ns.py 2>&1 | grep "LAST\|^finalize" | tail -4; fi
echo "== parity"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== stages"; AB_CONFIGS=${AB_CONFIGS:-c2plane,c2,c4} timeout 900 python tools/ab2.py ${AB_VARIANTS:-base} 2>&1 | tee gpurun_out/ab_quick.log
