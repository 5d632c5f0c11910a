"""Per-launch breakdown of one hand-written training step (C3 shape, 8192 rows): CUDA events around every C-ABI call of
train_engine.TrainEngine.step, summed by call site and shape.  Run on the B200 box."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from usflows_b200 import ops, train_engine
from usflows_b200.builders import build_flow
from oracle import flow_oracle as O
import bench

spec = bench.WORKLOADS["c2"]["spec"]
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
flow = build_flow(spec, O.random_params(spec, 0), device="cuda", precision="fp32")
x = torch.rand(rows, 784, device="cuda")
eng = train_engine.TrainEngine(flow, rows)
for _ in range(3):
    eng.step(x, rows)
torch.cuda.synchronize()
records = []
names = ["linear", "linear_splitk", "planes_glue", "mat_prep", "base_backward", "base_logprob", "tri_inverse_batched", "tri_mask",
         "lu_assemble", "ingest"]
orig = {n: getattr(ops, n) for n in names}


def wrap(name, f):
    def inner(*a, **k):
        if name == "linear":
            label = f"linear M={a[1].rows} N={a[4]} K={a[5]} eng={a[0]}" + (" +resid" if k.get("resid") is not None else "")
        elif name == "linear_splitk":
            label = f"splitk out={a[1].rows}x{a[3]} K={a[4]} split={a[6]}"
        elif name == "planes_glue":
            label = f"glue rows={k['rows']} n={k['n']}" + (" T" if k.get("t") is not None else "") + (" out" if k.get("out") is not None else "") + (" mask" if k.get("mask_h") is not None else "")
        elif name == "mat_prep":
            ref = k["out_f32"] if k.get("out_f32") is not None else k["out"].h16
            label = f"mat_prep {tuple(ref.shape)}"
        else:
            label = name
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(*a, **k); e1.record()
        records.append((label, e0, e1))
    return inner


for n, f in orig.items():
    setattr(ops, n, wrap(n, f))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); eng.step(x, rows); t1.record()
torch.cuda.synchronize()
for n, f in orig.items():
    setattr(ops, n, f)
agg = collections.OrderedDict()
for label, e0, e1 in records:
    ms = e0.elapsed_time(e1)
    c = agg.setdefault(label, [0, 0.0])
    c[0] += 1; c[1] += ms
print(f"step (instrumented): {t0.elapsed_time(t1):.3f} ms, {len(records)} C-ABI launches")
for label, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms  {n:4d} x {1e3 * ms / n:8.1f} us  {label}")
print(f"sum of launches: {sum(v[1] for v in agg.values()):.3f} ms")
