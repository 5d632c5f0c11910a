"""One dense and one triangular fp64 product at d = 3072 (the C5 layer width) for an `ncu --set full` capture of
`matmul_f64_mma_kernel` (run: ncu --set full --clock-control none -k regex:matmul_f64 -c 2 python tools/f64_ncu.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from usflows_b200 import ops

d = 3072
a = torch.randn(d, d, dtype=torch.float64, device="cuda")
b = torch.randn(d, d, dtype=torch.float64, device="cuda")
o = torch.empty_like(a)
ops.matmul_f64(a, b, o)
ops.matmul_f64(a.tril(), b.triu(), o, ops.TRI_LOWER_UPPER)
torch.cuda.synchronize()
