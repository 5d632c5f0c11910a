#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_training.py -m gpu -q -p no:cacheprovider -k "split_k or glue or base_backward or triangular or mat_prep or hand_written or train_step or gradients_match or fit_runs or captured" > gpurun_out/r2c_pytest_train.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c_pytest_train.log
tail -12 gpurun_out/r2c_pytest_train.log
timeout 300 python tools/train_breakdown.py > gpurun_out/r2c_train_breakdown.log 2>&1; cat gpurun_out/r2c_train_breakdown.log | head -60
timeout 300 python tools/train_profile.py > gpurun_out/r2c_train_profile.log 2>&1; grep "ms/step\|hand-written" gpurun_out/r2c_train_profile.log
