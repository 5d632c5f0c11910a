#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "row_shard" > gpurun_out/r2g_shard.log 2>&1
grep -v "^$" gpurun_out/r2g_shard.log | tail -60
