"""Cost of one weight-version change on the inference path (what a validation log_prob after an optimiser step pays):
weight preparation + launch-program rebuild (+ graph re-capture for small batches).  Run on the B200 box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from usflows_b200.builders import build_flow
from oracle import flow_oracle as O
import bench

for wl in ("c2", "c5"):
    spec = bench.WORKLOADS[wl]["spec"]
    flow = build_flow(spec, O.random_params(spec, 0), device="cuda", precision="fp32")
    d = spec["in_dims"][0]
    for rows in (256, 65536 if wl == "c2" else 32768):
        x = torch.rand(rows, d, device="cuda")
        flow.log_prob(x); flow.log_prob(x)
        torch.cuda.synchronize()
        ts = []
        for it in range(4):
            with torch.no_grad():
                for p in flow.parameters():
                    p.mul_(1.0 + 1e-4)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            flow.log_prob(x)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            flow.log_prob(x)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
        first = sorted(t[0] for t in ts)[len(ts) // 2]
        steady = sorted(t[1] for t in ts)[len(ts) // 2]
        print(f"{wl} rows {rows}: log_prob right after a weight update {first:.1f} ms, same weights again {steady:.2f} ms "
              f"-> weight preparation + program rebuild {first - steady:.1f} ms", flush=True)
