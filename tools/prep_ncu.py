"""One weight-version change followed by one small log_prob, bracketed by cudaProfilerStart/Stop -- the launch list of the
weight preparation (run under `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from usflows_b200.builders import build_flow
from oracle import flow_oracle as O
import bench

spec = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]["spec"]
flow = build_flow(spec, O.random_params(spec, 0), device="cuda", precision="fp32")
x = torch.rand(256, spec["in_dims"][0], device="cuda")
flow.log_prob(x); flow.log_prob(x)
with torch.no_grad():
    for p in flow.parameters():
        p.mul_(1.0 + 1e-4)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
flow.log_prob(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
