python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host or chunk" 2>&1 | tail -15 > gpurun_out/r1t_pytest.log
python - > gpurun_out/r1t_e2e.log 2>&1 <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from helpers import build_flow
from oracle import flow_oracle as O
import bench
from usflows_b200 import flows as F
spec = bench.WORKLOADS["c2"]["spec"]
flow = build_flow(spec, O.random_params(spec, 0), device="cuda", precision="fp32")
rows, d = 65536, 784
x_host = torch.rand(rows, d).pin_memory()
out_host = torch.empty(rows).pin_memory()
x = x_host.cuda()
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print(f"device log_prob: {timed(lambda: flow.log_prob(x)):.3f} ms")
for growth, mx in ((1, 1), (2, 2), (2, 4), (2, 8), (3, 9), (4, 4), (4, 16)):
    F.HOST_CHUNK_GROWTH, F.HOST_CHUNK_MAX_UNITS = growth, mx
    ms = timed(lambda: flow.log_prob_host(x_host, out_host))
    print(f"growth {growth} max {mx}: {ms:.3f} ms = {rows/ms/1e3:.2f} M rows/s")
PY
cat gpurun_out/r1t_pytest.log | tail -5; cat gpurun_out/r1t_e2e.log
