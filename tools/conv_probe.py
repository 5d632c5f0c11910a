"""Micro-benchmark of the implicit-GEMM convolution (usf_conv2d_rows) against gather + contraction (run on the B200 box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from usflows_b200 import _lib, engine, ops
from usflows_b200.ops import Act

n, H, W, C, N, k = 16384, 7, 7, 32, 32, 3
rows = n * H * W
g = torch.Generator().manual_seed(0)
x = torch.randn(rows, C, generator=g).cuda()
w = (torch.randn(N, k * k * C, generator=g) / 17).cuda()
b = torch.randn(N, generator=g).cuda()
w_hi, w_lo = engine._operand(w, "fp32_tf32", ops.ENGINE_TC_3XTF32)
wh16, wl16 = engine._operand(w, "fp32", ops.ENGINE_TC_3XF16)
out = Act(rows, N, f32=torch.empty(rows, N, device="cuda"))
cols = Act(rows, k * k * C, h16=torch.empty(rows, k * k * C, dtype=torch.float16, device="cuda"),
           l16=torch.empty(rows, k * k * C, dtype=torch.float16, device="cuda"))

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

print(f"rows {rows}, C {C}, N {N}, k {k}: {rows * 2 * k * k * C * N / 1e9:.1f} GFLOP")
for chunk in (1, 2, 3, 5, 9):
    _lib.load().usf_set_accum_chunk(chunk)
    us = timed(lambda: ops.conv2d_rows(x, n, H, W, C, k, 1, w_hi, w_lo, N, bias=b, relu=True, out=out, relu_in=True))
    print(f"conv2d_rows accum chunk {chunk}: {us:.1f} us")
_lib.load().usf_set_accum_chunk(2)
for flags, what in ((64, "gather without loads / stores (barriers only)"), (128, "no MMAs"), (192, "neither")):
    _lib.load().usf_debug_gemm_timeline(None, flags)
    us = timed(lambda: ops.conv2d_rows(x, n, H, W, C, k, 1, w_hi, w_lo, N, bias=b, relu=True, out=out, relu_in=True))
    print(f"conv2d_rows, {what}: {us:.1f} us")
_lib.load().usf_debug_gemm_timeline(None, 0)
us1 = timed(lambda: ops.im2col(x, n, H, W, C, k, 1, cols, relu=True))
us2 = timed(lambda: ops.linear(ops.ENGINE_TC_3XF16, cols, wh16, wl16, N, k * k * C, bias=b, relu=True, out=out))
print(f"im2col {us1:.1f} us + linear {us2:.1f} us = {us1 + us2:.1f} us")
for rr in (128 * 148, 128 * 148 * 4):
    nn = rr // 49
    us = timed(lambda: ops.conv2d_rows(x, nn, H, W, C, k, 1, w_hi, w_lo, N, bias=b, relu=True, out=out, relu_in=True))
    print(f"conv2d_rows {nn} images ({nn * 49 / 128 / 148:.2f} tiles per SM): {us:.1f} us")
