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

# ---- the pixel-plane route (usf_conv2d_pix): one 4-D TMA box per tap, no gather threads -------------------------------
from usflows_b200 import image_engine
a16 = torch.empty(rows, 64, dtype=torch.float16, device="cuda")
b16 = torch.empty(rows, 64, dtype=torch.float16, device="cuda")
y = torch.randn(rows, 32, generator=g).cuda()
x16c = torch.randn(rows, 16, generator=g).cuda()
mask = (torch.rand(49 * 16, generator=g) > 0.5).float().cuda()
w16 = image_engine._pix_weight(w, k * k, C, 32, None)
b32 = image_engine._pad_vec(b, 32)
w2 = (torch.randn(64, 32, generator=g) / 6).cuda()
w16_2 = image_engine._pix_weight(w2, 1, 32, 64, None)
b64 = torch.randn(64, generator=g).cuda()
gamma, beta = torch.ones(32, device="cuda"), torch.zeros(32, device="cuda")
us = timed(lambda: ops.pix_encode(x16c, 49, a16, mask=mask))
print(f"pix_encode (16 channels, mask): {us:.1f} us")
ops.pix_encode(x, 49, a16, relu=True)
for taps in (1, 2, 3):
    _lib.load().usf_set_pix_chain_taps(taps)
    us = timed(lambda: ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 32, out_f32=out.f32, out16=b16, relu_planes=True))
    print(f"conv2d_pix plain (fp32 rows + planes out), {taps} taps per chain: {us:.1f} us")
    us = timed(lambda: ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 32, gated=True, post_relu=True, w2=w16_2, bias2=b64,
                                      gamma=gamma, beta=beta, eps=1e-5, out_f32=y, out16=b16, relu_planes=True))
    print(f"conv2d_pix gated block (conv 3x3 + conv 1x1 + gate + ReLU + LayerNorm), {taps} taps per chain: {us:.1f} us")
_lib.load().usf_set_pix_chain_taps(0)
us = timed(lambda: ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 16, x=x16c, inv_mask=mask, sign=1.0))
print(f"conv2d_pix last (coupling update of 16 channels): {us:.1f} us")
for nn in (5 * 148, 5 * 148 * 4):
    us = timed(lambda: ops.conv2d_pix(a16, nn, H, W, k, 1, w16, b32, 32, out_f32=out.f32, out16=b16))
    print(f"conv2d_pix plain {nn} images ({nn / 5 / 148:.0f} tiles per SM): {us:.1f} us")

for flags, what in ((64, "no activation loads"), (128, "no MMAs"), (256, "no epilogue loads / stores"), (64 | 128, "no loads, no MMAs"),
                    (64 | 256, "no loads, no epilogue traffic"), (64 | 128 | 256, "skeleton only")):
    _lib.load().usf_debug_gemm_timeline(None, flags)
    us = timed(lambda: ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 32, out_f32=out.f32, out16=b16, relu_planes=True))
    us2 = timed(lambda: ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 32, gated=True, post_relu=True, w2=w16_2, bias2=b64,
                                       gamma=gamma, beta=beta, eps=1e-5, out_f32=y, out16=b16, relu_planes=True))
    print(f"conv2d_pix, {what}: plain {us:.1f} us, gated {us2:.1f} us")
_lib.load().usf_debug_gemm_timeline(None, 0)
us = timed(lambda: ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 32, out16=b16, relu_planes=True))
print(f"conv2d_pix plain, planes out only: {us:.1f} us")
us = timed(lambda: ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 32, out_f32=out.f32))
print(f"conv2d_pix plain, fp32 rows out only: {us:.1f} us")


# time against tiles per SM (5 images per tile, 148 SMs)
for flags in (0, 64 | 256, 64 | 128 | 256):
    _lib.load().usf_debug_gemm_timeline(None, flags)
    line = []
    for tps in (1, 2, 3, 4, 6, 8, 12, 16, 22):
        nn = 5 * 148 * tps
        us = timed(lambda: ops.conv2d_pix(a16, nn, H, W, k, 1, w16, b32, 32, out_f32=out.f32, out16=b16, relu_planes=True))
        line.append(f"{tps}: {us:.1f}")
    print(f"conv2d_pix plain, flags {flags}, us by tiles per SM: " + ", ".join(line))
_lib.load().usf_debug_gemm_timeline(None, 0)

for ga in (0, 1, 2, 3):
    _lib.load().usf_set_pix_gate_at(ga)
    line = []
    for taps in (1, 2, 3):
        _lib.load().usf_set_pix_chain_taps(taps)
        us = timed(lambda: ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 32, gated=True, post_relu=True, w2=w16_2, bias2=b64,
                                          gamma=gamma, beta=beta, eps=1e-5, out_f32=y, out16=b16, relu_planes=True))
        line.append(f"{taps} taps/chain {us:.1f} us")
    print(f"conv2d_pix gated, gate contraction behind {ga} chains of the next tile: " + ", ".join(line))
_lib.load().usf_set_pix_gate_at(1)
_lib.load().usf_set_pix_chain_taps(0)
