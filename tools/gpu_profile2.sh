#!/bin/bash
# ncu passes of round 2 (B200_PROFILING.md): each ncu command runs only after the identical plain command exited 0.
set -u
mkdir -p gpurun_out
A="python tools/prof_target.py c2plane 6"
$A > gpurun_out/prof_plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2plane.csv $A > gpurun_out/ncu_a1.log 2>&1
echo "launch list rc=$?"
$A > gpurun_out/prof_plain_a2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -c 5 -o gpurun_out/r2_knn_c2plane $A > gpurun_out/ncu_a2.log 2>&1
echo "knn (normals epilogue k=20 + k=1 iterations) rc=$?"
$A > gpurun_out/prof_plain_a3.log 2>&1 &&
ncu --set full --clock-control none -k regex:"hist_kernel|accumulate_kernel" -s 6 -c 4 -o gpurun_out/r2_select_minimize $A > gpurun_out/ncu_a3.log 2>&1
echo "select/minimise rc=$?"
B="python tools/prof_target.py c4 3"
$B > gpurun_out/prof_plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"knn_kernel|knn_overflow" -s 2 -c 4 -o gpurun_out/r2_knn_c4 $B > gpurun_out/ncu_b.log 2>&1
echo "knn k=10 rc=$?"
C="python tools/prof_target.py normals10m"
$C > gpurun_out/prof_plain_c.log 2>&1 &&
ncu --set full --clock-control none -k regex:"knn_kernel|knn_overflow" -c 2 -o gpurun_out/r2_normals_10m $C > gpurun_out/ncu_c.log 2>&1
echo "normals 10M rc=$?"
ls -la gpurun_out | tail -12
