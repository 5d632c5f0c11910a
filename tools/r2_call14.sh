#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x -k "half_size or split_k or tma_store_path_equals or all_engines" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2o_pytest.log
tail -6 gpurun_out/r2o_pytest.log
for slab in 128 64; do for s in 65536x784x784 65536x1024x1024 65536x392x1024 65536x1024x392; do python tools/gemm_timeline.py --engine 3xf16 --shape $s --flags 0 --slab $slab 2>&1 | grep engine | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($slab, d['N'], d['K'], 'us', d['us'], 'period', d.get('period'), 'wait_operands', d.get('wait_empty_to_full'), 'store', d.get('store'))"; done; done
