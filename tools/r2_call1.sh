#!/bin/bash
# round-2 GPU call 1: test suite, default bench line, training profile, compute-sanitizer
mkdir -p gpurun_out
rm -f gpurun_out/parity_elementwise.jsonl
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc $?"
tail -c 600 gpurun_out/r2a_bench.err
timeout 300 python tools/train_profile.py > gpurun_out/r2a_train_profile.log 2>&1; tail -30 gpurun_out/r2a_train_profile.log
timeout 300 python bench.py --workload c1 --no-modes --no-extra --no-train > gpurun_out/r2a_c1.json 2> gpurun_out/r2a_c1.err
bash tools/sanitize.sh memcheck synccheck racecheck
