#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2l_bench8.json 2> gpurun_out/r2l_bench8.err; echo "rc $?"
tail -3 gpurun_out/r2l_bench8.err
nvidia-smi topo -m > gpurun_out/r2l_topo.txt 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench8.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('train', {k:d['train'][k] for k in ('value','ms_per_step','allreduce_ms_alone','global_batch')})
print('h2d', d['h2d'])
print('affinity', d['config']['host_affinity'])
for e in d['configs']['c5_sweep']['entries']: print({k:(round(x,3) if isinstance(x,float) else x) for k,x in e.items() if k in ('points','precision','log_prob_samples_per_sec','sample_samples_per_sec','frac','skipped')})
PY
