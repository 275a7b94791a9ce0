#!/bin/bash
# ncu passes, end of round 2 (B200_PROFILING.md): each ncu command runs only after the identical plain command exited 0.
set -u
mkdir -p gpurun_out
BN="python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu --no-extra --no-e2e"
timeout 600 $BN > gpurun_out/r2e_bench_plain.json 2> gpurun_out/r2e_bench_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_end_launches_bench.csv $BN > gpurun_out/ncu_e0.log 2>&1
echo "bench launch list rc=$?"
A="python tools/prof_target.py c2plane 8"
timeout 300 $A > gpurun_out/prof_plain_e1.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"knn_kernel|knn_overflow" -s 8 -c 6 -o gpurun_out/r2_end_knn_c2plane $A > gpurun_out/ncu_e1.log 2>&1
echo "knn rc=$?"
timeout 300 $A > gpurun_out/prof_plain_e2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_accumulate|finalize_kernel" -s 6 -c 6 -o gpurun_out/r2_end_select_accumulate $A > gpurun_out/ncu_e2.log 2>&1
echo "select_accumulate rc=$?"
tail -3 gpurun_out/ncu_e2.log
ls -la gpurun_out | tail -8
