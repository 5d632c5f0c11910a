"""usf_matmul_f64 against cuBLAS DGEMM (torch.matmul) at the operator sizes of the C2 / C5 stacks -- what the fp64
composition of neighbouring affine layers costs per weight version, and the rate the library DGEMM sets as the bar."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from usflows_b200 import ops

def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3

for d in (784, 1024, 3072):
    a = torch.randn(d, d, dtype=torch.float64, device="cuda")
    b = torch.randn(d, d, dtype=torch.float64, device="cuda")
    o = torch.empty_like(a)
    t_own = timed(lambda: ops.matmul_f64(a, b, o))
    t_lib = timed(lambda: torch.matmul(a, b, out=o))
    fl = 2.0 * d ** 3
    lo, up = a.tril(), b.triu()
    t_lu = timed(lambda: ops.matmul_f64(lo, up, o, ops.TRI_LOWER_UPPER))
    t_ul = timed(lambda: ops.matmul_f64(up, lo, o, ops.TRI_UPPER_LOWER))
    print(f"d={d}: lower x upper {t_lu:8.1f} us, upper x lower {t_ul:8.1f} us ({t_lu / t_own:.2f} / {t_ul / t_own:.2f} of the dense product)")
    print(f"d={d}: usf_matmul_f64 {t_own:8.1f} us ({fl / t_own * 1e-6:6.2f} TFLOP/s)   cuBLAS DGEMM {t_lib:8.1f} us "
          f"({fl / t_lib * 1e-6:6.2f} TFLOP/s)", flush=True)
