#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x -k "tma_store or staged_store or all_engines" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2n_pytest.log
tail -8 gpurun_out/r2n_pytest.log
for s in 65536x1024x1024 65536x784x784; do python tools/gemm_timeline.py --engine bf16 --shape $s --flags 0,2 2>&1 | grep engine | cut -c1-140; done
timeout 300 python bench.py --no-extra --no-train --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown_ms'])
print({m:(round(v['value']/1e6,2), round(v['frac'],3), round(v['frac_of_mode_peak'],3)) for m,v in d['modes'].items()})
PY
