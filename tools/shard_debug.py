import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_case, build_flow
from usflows_b200 import parallel
spec, params, arr = load_case("d100_h50_hh")
g = torch.Generator().manual_seed(12)
x = torch.rand(30001, 100, generator=g).pin_memory()

def halves(got, want, tag):
    n = got.shape[0] // 2
    print(tag, "first half equal:", torch.equal(got[:n], want[:n]), "second:", torch.equal(got[n:], want[n:]),
          "second sample", got[n:n + 2].tolist(), want[n:n + 2].tolist(), flush=True)

flow = build_flow(spec, params)
want = flow.log_prob(x.cuda()).cpu()
sf = parallel.ShardedFlow(flow)
# 1. the replica alone, from the main thread, on its shard
rep = sf.replicas[1]
n = 15000
out = torch.empty(30001 - n, pin_memory=True)
rep.log_prob_host(x[n:], out)
print("replica alone (main thread):", torch.equal(out, want[n:]), flush=True)
# 2. through the thread pool, one after the other
import concurrent.futures as cf
with cf.ThreadPoolExecutor(1) as pool:
    out2 = torch.empty(30001 - n, pin_memory=True)
    pool.submit(rep.log_prob_host, x[n:], out2).result()
    print("replica alone (worker thread):", torch.equal(out2, want[n:]), flush=True)
# 3. both at once
halves(sf.log_prob(x), want, "sharded call 1")
halves(sf.log_prob(x), want, "sharded call 2")
with torch.no_grad():
    for p in flow.parameters():
        p.mul_(1.0 + 1e-3)
want = flow.log_prob(x.cuda()).cpu()
halves(sf.log_prob(x), want, "after weight update")
# 4. a fresh sharded flow whose first use is concurrent
flow2 = build_flow(spec, params)
want2 = flow2.log_prob(x.cuda()).cpu()
halves(parallel.ShardedFlow(flow2).log_prob(x), want2, "fresh, concurrent first use")
