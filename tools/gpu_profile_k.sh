#!/bin/bash
# full ncu capture of arbitrary kernels of the bench command: tools/gpu_profile_k.sh <regex> <skip> <count> <outname>
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps ${PROF_STEPS:-12} --warmup 3 --no-extra --no-cpu --no-e2e"
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/$4 $CMD > gpurun_out/ncu_full_$4.log 2>&1
echo "capture rc=$?"
