"""Timing probe: coupling-output contraction (in-place fp16-split residual) vs the same contraction without residual."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from usflows_b200 import ops
from usflows_b200.ops import Act
from gemm_bench import make_case

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

M, N, K = 65536, 392, 1024
act, wt, wl, bias, out, _, _ = make_case("3xf16", M, N, K, 0, True)
# a [M, 784] stream; the coupling writes its second half in place
ld = 784
sh, sl = torch.zeros(2, M, ld, dtype=torch.float16, device="cuda")
seg = Act(M, N, h16=sh[:, 392:], l16=sl[:, 392:])
print("no residual, separate out   : %.1f us" % timed(lambda: ops.linear(ops.ENGINE_TC_3XF16, act, wt, wl, N, K, bias=bias, out=out)))
print("no residual, out = segment  : %.1f us" % timed(lambda: ops.linear(ops.ENGINE_TC_3XF16, act, wt, wl, N, K, bias=bias, out=seg)))
print("residual in place (coupling): %.1f us" % timed(lambda: ops.linear(ops.ENGINE_TC_3XF16, act, wt, wl, N, K, bias=bias, resid=seg, resid_sign=-1.0, out=seg)))
