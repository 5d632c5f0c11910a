"""log_prob latency at small batch sizes with / without CUDA-graph replay (run on the B200 box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from usflows_b200 import flows
from usflows_b200.builders import build_flow
from oracle import flow_oracle as O
import bench
for wl in ("c2", "c5"):
    spec = bench.WORKLOADS[wl]["spec"]
    flow = build_flow(spec, O.random_params(spec, 0), device="cuda", precision="fp32")
    d = spec["in_dims"][0]
    for rows in (256, 1024, 4096, 16384):
        x = torch.rand(rows, d, device="cuda")
        res = {}
        for graphs in (0, 16384):
            flows.SMALL_BATCH_GRAPH_ROWS = graphs
            for _ in range(3): flow.log_prob(x)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(20): lp = flow.log_prob(x)
            torch.cuda.synchronize(); res[graphs] = (time.perf_counter() - t0) / 20 * 1e3
            ref = lp if graphs == 0 else ref
        same = bool(torch.equal(ref, lp))
        print(f"{wl} rows {rows}: launch-by-launch {res[0]:.3f} ms, graph replay {res[16384]:.3f} ms, identical {same}")
