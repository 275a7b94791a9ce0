#!/bin/bash
# sharded loop A/B over run-time switches on N GPUs: tools/gpu_multi_ab.sh N "" "PMGPU_DEFER_FINALIZE=0" ...
set -u
N=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  for cfg in c2plane c4; do
    env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus $N --config $cfg --no-extra --no-cpu --no-e2e --reps 5 2> gpurun_out/mab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=$N %-28s %-8s value %8.1f it/s  stages %s' % ('[$v]', '$cfg', d['value'], {k: round(v,4) for k,v in d['extra']['stage_ms_per_iteration'].items() if v}))" || tail -5 gpurun_out/mab.err
  done
done
