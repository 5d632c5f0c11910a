"""Does a smaller, wave-aligned row chunk (activations of a layer stay in the 126 MB L2) beat one 65 536-row chunk on the C2
log_prob step?  Same rows, same kernels; only the chunking differs.  Run on the B200 box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import usflows_b200 as U
from usflows_b200.builders import build_flow
from oracle import flow_oracle as O
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
spec = bench.WORKLOADS["c2"]["spec"]
flow = build_flow(spec, O.random_params(spec, 0), device="cuda", precision=prec)
unit = 74 * 256
for units_total, chunk_units in ((3, 3), (3, 1), (4, 4), (4, 2), (4, 1), (6, 6), (6, 3), (6, 2)):
    rows = unit * units_total
    x = torch.rand(rows, 784, device="cuda")
    U.set_chunk_rows(unit * chunk_units)
    for _ in range(3):
        flow.log_prob(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        flow.log_prob(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{prec} rows {rows} in chunks of {unit * chunk_units}: {ms:.3f} ms  {rows / ms / 1e3:.2f} M rows/s", flush=True)
