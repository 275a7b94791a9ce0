#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== knn parity tests on the new build"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_ab.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_ab.log
echo "== A/B"; timeout 1500 python tools/ab2.py "$@" 2>&1 | tee gpurun_out/ab2.log
