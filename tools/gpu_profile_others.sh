#!/bin/bash
# ncu --set full of the kernels other than K2 (select, minimise, build, stage 2), a handful of launches each
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --no-extra --no-cpu --no-e2e"
$CMD > gpurun_out/prof_plain4.log 2>&1 &&
ncu --set full --clock-control none -k regex:"hist_kernel|accumulate_kernel|subtree_kernel|knn_overflow_kernel|level_boxes_kernel" -s 14 -c 22 -o gpurun_out/prof_others $CMD > gpurun_out/ncu_full_others.log 2>&1
echo "capture rc=$?"; ls -la gpurun_out/prof_others.ncu-rep
