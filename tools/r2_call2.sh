#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_training.py -m gpu -q -p no:cacheprovider -k "split_k or glue or base_backward or triangular or mat_prep or hand_written or train_step or gradients_match or fit_runs" > gpurun_out/r2b_pytest_train.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest_train.log
tail -25 gpurun_out/r2b_pytest_train.log
timeout 300 python tools/train_profile.py > gpurun_out/r2b_train_profile.log 2>&1; tail -32 gpurun_out/r2b_train_profile.log
python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
