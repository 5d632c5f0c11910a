cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/conv_probe.py > gpurun_out/conv_probe.log 2>&1; grep "by tiles" gpurun_out/conv_probe.log | cut -c1-600
