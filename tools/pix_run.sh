cd /root/repo
timeout 900 python -m pytest tests/test_gpu_conv_pix.py -x -q -m gpu -k "falls_back" 2>&1 | tail -12
