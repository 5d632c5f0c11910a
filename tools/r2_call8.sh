#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_elementwise.jsonl
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2j_pytest.log
tail -12 gpurun_out/r2j_pytest.log
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'frac', d['roofline']['frac'])
print('train', d['train']['ms_per_step'], d['train']['value'])
print(d['breakdown_ms'])
for e in d['configs']['c5_sweep']['entries'][:3]: print({k:(round(x,3) if isinstance(x,float) else x) for k,x in e.items() if k in ('points','precision','log_prob_samples_per_sec','log_prob_ms','frac')})
print({m:v['value'] for m,v in d['modes'].items()})
PY
python tools/small_batch_probe.py 2>&1 | tail -8
