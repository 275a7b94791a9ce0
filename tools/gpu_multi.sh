#!/bin/bash
# A multi-GPU box visit (gpurun --gpus N): sharded-vs-single parity on 2 GPUs, then the bench at N ranks.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_multi.txt 2>&1
echo "== sharded parity (2 ranks)"; timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/pytest_sharded.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 tests/sharded_worker.py --points 1000000 --json gpurun_out/sharded_parity_1M.json > gpurun_out/sharded_1M.log 2>&1; echo "1M parity rc=$?"; tail -3 gpurun_out/sharded_1M.log | cut -c1-1500
for n in ${SCALE_NS:-$N}; do
echo "== bench N=$n"
if [ "$n" = "1" ]; then
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ${BENCH_ARGS:-} > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
else
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2975$n bench.py --gpus $n --steps 20 --warmup 5 ${BENCH_ARGS:-} > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
fi
echo "rc=$?"; tail -4 gpurun_out/bench_n$n.err | cut -c1-600; cut -c1-2500 gpurun_out/bench_n$n.json
done
