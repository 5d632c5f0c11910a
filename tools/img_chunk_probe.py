"""Image step time against the chunk size (channels-last rows per chunk): L2 residency of a chunk's buffers vs launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_flow
from oracle import flow_oracle as O
from usflows_b200 import image_engine

spec = dict(in_dims=[16, 7, 7], coupling_blocks=15, conditioner="convnet2d", c_hidden=32, num_layers=3, kernel_size=3,
            gating=True, normalize_layers=True, affine_conjugation=True, lu_transform=1, householder=0, base="radial",
            p=1, norm="lognormal")
flow = build_flow(spec, O.random_params(spec, 0))
x = torch.rand(16384, 16, 7, 7, device="cuda")
for rows in (1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21):
    image_engine.IMAGE_CHUNK_ROWS = image_engine.IMAGE_CHUNK_ROWS_PIX = rows
    for _ in range(3): flow.log_prob(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): flow.log_prob(x)
    e1.record(); torch.cuda.synchronize()
    print(f"chunk {rows} rows ({rows // 49} images): {e0.elapsed_time(e1) / 5:.2f} ms per 16 384-image log_prob")
