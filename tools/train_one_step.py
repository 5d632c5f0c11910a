"""One hand-written training step (C3 shape, 8192 rows) inside a cudaProfilerStart/Stop range, launch by launch (no graph
replay), for `ncu --profile-from-start off`.  Run on the B200 box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import usflows_b200 as U
from usflows_b200 import training
from usflows_b200.builders import build_flow
from oracle import flow_oracle as O
import bench
spec = bench.WORKLOADS["c2"]["spec"]
flow = build_flow(spec, O.random_params(spec, 0), device="cuda", precision="fp32")
opt = U.SophiaG(list(flow.parameters()), lr=1e-3, weight_decay=0.0)
ts = training.TrainStep(flow, opt, distributed=False, graph=False)
x = torch.rand(8192, 784, device="cuda")
for _ in range(2):
    ts.step(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ts.step(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done, hand-written pass:", ts.use_engine)
