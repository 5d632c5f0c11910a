#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "tma_store_path or staged_store or mat_prep" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
for f in 0 256; do python tools/gemm_bench.py --engines 3xf16 --shapes 65536x392x1024,65536x1536x1024 --resid --flags $f --iters 10 2>&1 | grep engine; done
python tools/gemm_bench.py --engines 3xf16 --shapes 65536x392x1024,65536x784x784,65536x1024x1024 --iters 10 2>&1 | grep engine
timeout 300 python bench.py --no-extra --no-train --no-modes --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown_ms'])
PY
timeout 300 python tools/train_profile.py > gpurun_out/r2e_train_profile.log 2>&1; grep "ms/step\|hand-written" gpurun_out/r2e_train_profile.log
