#!/bin/bash
# End-of-round visit: smoke, bench (both arms, default flags), then the other BASELINE configs.  Logs go to gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2e_bench_ref.json
echo "== bench ours"; timeout 1500 python bench.py > gpurun_out/r2e_bench_ours.json 2> gpurun_out/r2e_bench_ours.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench_ours.err; cut -c1-1200 gpurun_out/r2e_bench_ours.json
for c in c3 c4 c5; do
echo "== bench $c"; timeout 1200 python bench.py --config $c > gpurun_out/r2e_bench_$c.json 2> gpurun_out/r2e_bench_$c.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2e_bench_$c.json
done
