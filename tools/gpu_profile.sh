#!/bin/bash
# ncu passes (B200_PROFILING.md): launch list of the bench command + one full capture of the
# dominant kernel.  Each ncu command runs only after the identical plain command exited 0.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps ${PROF_STEPS:-4} --warmup 3 --no-extra --no-cpu --no-e2e"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s ${PROF_SKIP:-4} -c 2 -o gpurun_out/prof_knn $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -12
