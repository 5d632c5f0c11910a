"""One usf_conv2d_pix launch (plain or gated) for an ncu capture: python tools/pix_ncu.py [gated]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from usflows_b200 import image_engine, ops

gated = len(sys.argv) > 1 and sys.argv[1] == "gated"
n, H, W, k = 16384, 7, 7, 3
rows = n * H * W
g = torch.Generator().manual_seed(0)
x = torch.randn(rows, 32, generator=g).cuda()
w = (torch.randn(32, k * k * 32, generator=g) / 17).cuda()
b = torch.randn(32, generator=g).cuda()
a16 = torch.empty(rows, 64, dtype=torch.float16, device="cuda")
b16 = torch.empty(rows, 64, dtype=torch.float16, device="cuda")
y = torch.randn(rows, 32, generator=g).cuda()
out = torch.empty(rows, 32, device="cuda")
w16, b32 = image_engine._pix_weight(w, k * k, 32, 32, None), image_engine._pad_vec(b, 32)
w16_2 = image_engine._pix_weight((torch.randn(64, 32, generator=g) / 6).cuda(), 1, 32, 64, None)
b64 = torch.randn(64, generator=g).cuda()
gamma, beta = torch.ones(32, device="cuda"), torch.zeros(32, device="cuda")
ops.pix_encode(x, 49, a16, relu=True)
for _ in range(3):
    if gated:
        ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 32, gated=True, post_relu=True, w2=w16_2, bias2=b64, gamma=gamma, beta=beta,
                       eps=1e-5, out_f32=y, out16=b16, relu_planes=True)
    else:
        ops.conv2d_pix(a16, n, H, W, k, 1, w16, b32, 32, out_f32=out, out16=b16, relu_planes=True)
torch.cuda.synchronize()
