#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 8 --steps 20 --warmup 3 --no-extra --no-modes --no-cpu-baseline > gpurun_out/r2m_train8_$tag.json 2> gpurun_out/r2m_train8_$tag.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2m_train8_$tag.json').read().strip().splitlines()[-1])
print('$tag', 'train ms', round(d['train']['ms_per_step'],3), 'allreduce alone', round(d['train']['allreduce_ms_alone'],3), 'log_prob ms', round(d['ms_per_step'],3))
PY
}
run default NCCL_DEBUG=WARN
run ch4 NCCL_MAX_NCHANNELS=4
run ch2 NCCL_MAX_NCHANNELS=2
run ch8cta NCCL_MAX_NCHANNELS=8
