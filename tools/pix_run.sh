cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv_pix.py -x -q -m gpu -k "narrow or chain" 2>&1 | tail -3
python tools/narrow_probe.py
