"""Timing of the SIMT engine's narrow fast path (16 x 16 1x1 convolution over 802 816 channels-last rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from usflows_b200 import ops
from usflows_b200.ops import Act
M = 16384 * 49
g = torch.Generator().manual_seed(0)
a = torch.randn(M, 16, generator=g).cuda()
w = torch.randn(16, 16, generator=g).cuda()
b = torch.randn(16, generator=g).cuda()
out = torch.empty(M, 16, device="cuda")
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
us = timed(lambda: ops.linear(0, Act(M, 16, f32=a), w, None, 16, 16, bias=b, out=Act(M, 16, f32=out)))
print(f"16 x 16 contraction over {M} rows: {us:.1f} us = {M * 128 / us / 1e6:.2f} TB/s of 64 B in + 64 B out per row")
