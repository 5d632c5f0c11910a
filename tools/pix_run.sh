cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_pix.py tests/test_image_flows.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/conv_probe.py > gpurun_out/conv_probe.log 2>&1; grep "conv2d_pix" gpurun_out/conv_probe.log | head -18 | cut -c1-200
timeout 600 python bench.py --workload mnist_img --no-train --no-extra > gpurun_out/bench_img.json 2> gpurun_out/bench_img.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_img.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('breakdown_ms'), d['gpu_launches'])
PY
