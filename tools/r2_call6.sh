#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err; echo "rc $?"
tail -5 gpurun_out/r2f_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('train', d['train'])
print('h2d', d['h2d'])
print('c4', {m:(v['log_prob_samples_per_sec'], v['sample_samples_per_sec']) for m,v in d['configs']['c4']['modes'].items()})
for e in d['configs']['c5_sweep']['entries']: print({k:(round(x,3) if isinstance(x,float) else x) for k,x in e.items() if k in ('points','precision','log_prob_samples_per_sec','frac','skipped')})
PY
python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "row_shard" 2>&1 | tail -3
