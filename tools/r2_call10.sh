#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_elementwise.jsonl
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2k_pytest.log
tail -6 gpurun_out/r2k_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2k_smoke.log 2>&1; tail -3 gpurun_out/r2k_smoke.log
bash tools/sanitize.sh memcheck racecheck synccheck
