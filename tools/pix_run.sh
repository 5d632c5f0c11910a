cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shard or replica or multi" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extra > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2f_bench_n2.json') if l.startswith('{')][-1])
print(d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'train', d.get('train',{}).get('value'), d.get('train',{}).get('ms_per_step'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload mnist_img --no-extra --no-train --no-modes > gpurun_out/r2f_bench_img_n2.json 2> gpurun_out/r2f_bench_img_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2f_bench_img_n2.json') if l.startswith('{')][-1])
print('img', d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
PY
