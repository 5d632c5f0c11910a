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
        ref = None
        for tag, graphs, plan in (("python", 0, False), ("c_plan", 0, True), ("graph", 16384, True)):
            flows.SMALL_BATCH_GRAPH_ROWS, flows.USE_C_PLAN = graphs, plan
            for _ in range(3): flow.log_prob(x)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(20): lp = flow.log_prob(x)
            torch.cuda.synchronize(); res[tag] = (time.perf_counter() - t0) / 20 * 1e3
            same = True if ref is None else bool(torch.equal(ref, lp))
            ref = lp if ref is None else ref
        print(f"{wl} rows {rows}: launch-by-launch from Python {res['python']:.3f} ms, one C call (usf_flow_logprob) "
              f"{res['c_plan']:.3f} ms, graph replay {res['graph']:.3f} ms, identical {same}")
