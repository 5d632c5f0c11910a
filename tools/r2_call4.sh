#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_training.py -m gpu -q -p no:cacheprovider -k "mat_prep or hand_written or train_step or captured" > gpurun_out/r2d_pytest_train.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2d_pytest_train.log
tail -5 gpurun_out/r2d_pytest_train.log
timeout 300 python tools/train_breakdown.py > gpurun_out/r2d_train_breakdown.log 2>&1; head -14 gpurun_out/r2d_train_breakdown.log
timeout 300 python tools/train_profile.py > gpurun_out/r2d_train_profile.log 2>&1; grep "ms/step\|hand-written" gpurun_out/r2d_train_profile.log; grep -A14 "Self CUDA %" gpurun_out/r2d_train_profile.log | cut -c1-60,150-230 | head -16
