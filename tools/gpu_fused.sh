#!/bin/bash
# fused select + minimise kernel: parity tests first, then A/B against the separate kernels
set -u
mkdir -p gpurun_out
echo "== probe"; PMGPU_VARIANT=prof timeout 200 python tools/probe_fused_ns.py 2>&1 | grep "select_accumulate" | tail -16
echo "== parity"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== A/B"; timeout 900 python tools/ab2.py base env:PMGPU_NO_FUSED_SELECT=1 2>&1 | tee gpurun_out/ab_fused.log
