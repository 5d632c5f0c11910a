"""Kernel-time profile of one MLE training step (C3 shape, 8192 rows) -- run on the B200 box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import usflows_b200 as U
from usflows_b200 import training
from usflows_b200.builders import build_flow
from oracle import flow_oracle as O
import bench
spec = bench.WORKLOADS["c2"]["spec"]
flow = build_flow(spec, O.random_params(spec, 0), device="cuda", precision="fp32")
params = list(flow.parameters())
opt = U.SophiaG(params, lr=1e-3, weight_decay=0.0)
x = torch.rand(8192, 784, device="cuda")
ts = training.TrainStep(flow, opt, distributed=False, engine=None if "--autograd" not in sys.argv else False)
print("hand-written pass:", ts.use_engine)
def step():
    ts.step(x)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / 5)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
