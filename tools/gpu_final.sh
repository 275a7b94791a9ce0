#!/bin/bash
# final visit of a round: the whole GPU suite, smoke, both bench arms with default flags, the other configs
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
bash tools/gpu_round3.sh 2>&1 | grep -v "^{"
