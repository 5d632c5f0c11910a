python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host or chunk" 2>&1 | tail -15 > gpurun_out/r1i_pytest.log
python - > gpurun_out/r1i_e2e.log 2>&1 <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from usflows_b200.builders import build_flow
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
for units in ((1,), (1, 2), (1, 1, 2), (1, 1, 2, 3), (1, 1, 2, 2), (1, 1, 1, 2, 2), (1, 2, 4)):
    F.HOST_CHUNK_UNITS = units
    ms = timed(lambda: flow.log_prob_host(x_host, out_host))
    print(f"chunk units {units}: {ms:.3f} ms = {rows/ms/1e3:.2f} M rows/s")
F.HOST_CHUNK_UNITS = (1, 1, 2, 3)
F.HOST_STAGE_BYTES = 0
ms = timed(lambda: flow.log_prob_host(x_host, out_host))
print(f"two-buffer ring, units (1, 1, 2, 3): {ms:.3f} ms = {rows/ms/1e3:.2f} M rows/s")
PY
cat gpurun_out/r1i_pytest.log | tail -5; cat gpurun_out/r1i_e2e.log
