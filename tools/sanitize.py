"""Small-shape pass over every kernel of libusflows_b200.so for compute-sanitizer (memcheck / racecheck /
synccheck over every kernel at small shapes).  Run on the GPU box through tools/sanitize.sh; the logs are summarised
under profiles/.  Shapes are tiny (a few tiles per kernel) because the sanitizer serialises and instruments everything."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flow_oracle as O  # noqa: E402  (parameters + expected values: this is a checker, like smoke())
from usflows_b200.builders import build_flow  # noqa: E402


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1.0))


def main():
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(3)
    # flat flow, every contraction engine (pair kernel: 2 M-blocks x 1-2 N-blocks, K tail), ingest, base density, sampling
    spec = dict(in_dims=[96], coupling_blocks=1, hidden_dims=[272, 128], affine_conjugation=True, lu_transform=1,
                householder=1, base="laplace")
    params = O.random_params(spec, 7)
    x = torch.rand(300, 96, generator=g)
    want = O.flow_log_prob(x, spec, params)
    for mode, tol in (("fp32", 2e-5), ("fp32_tf32", 2e-5), ("fp32_simt", 2e-5), ("tf32", 5e-2), ("bf16", 2e-1)):
        flow = build_flow(spec, params, device="cuda:0", precision=mode)
        lp = flow.log_prob(x.cuda())
        flow.backward(x.cuda())
        flow.sample([40])
        flow.log_prob_host(x.pin_memory())
        torch.cuda.synchronize()
        print(f"[{mode}] log_prob err {rel_err(lp, want):.2e}", flush=True)
        assert rel_err(lp, want) < tol
    # 784-wide layer shapes of the headline configuration on a few hundred rows (208-column tiles, 392-wide coupling halves)
    spec_w = dict(in_dims=[784], coupling_blocks=1, hidden_dims=[1024, 1024], affine_conjugation=True, lu_transform=1,
                  householder=0, base="laplace")
    pw = O.random_params(spec_w, 0)
    xw = torch.rand(520, 784, generator=g)
    fw = build_flow(spec_w, pw, device="cuda:0", precision="fp32")
    e = rel_err(fw.log_prob(xw.cuda()), O.flow_log_prob(xw, spec_w, pw))
    print(f"[c2 shapes] log_prob err {e:.2e}", flush=True)
    assert e < 2e-5
    # whole-flow kernel for tiny events
    spec1 = dict(in_dims=[2], coupling_blocks=3, hidden_dims=[32, 32], affine_conjugation=True, lu_transform=1,
                 householder=0, base="laplace")
    p1 = O.random_params(spec1, 5)
    x1 = torch.rand(700, 2, generator=g)
    f1 = build_flow(spec1, p1, device="cuda:0", precision="fp32")
    assert rel_err(f1.log_prob(x1.cuda()), O.flow_log_prob(x1, spec1, p1)) < 2e-5
    # ConvNet conditioner + radial base (gate_norm, radial kernels), image-shaped flow (layout, im2col, conv, masked add)
    spec2 = dict(in_dims=[64], coupling_blocks=1, conditioner="convnet", c_hidden=[64, 64], gating=True,
                 normalize_layers=True, affine_conjugation=True, lu_transform=1, householder=0, base="radial", p=2,
                 norm="gammamm", n_comp=5)
    spec3 = dict(in_dims=[16, 7, 7], coupling_blocks=1, conditioner="convnet2d", c_hidden=32, num_layers=1, kernel_size=3,
                 gating=True, normalize_layers=True, affine_conjugation=True, lu_transform=1, householder=0, base="radial",
                 p=1, norm="lognormal")
    from usflows_b200 import image_engine
    for tag, sp, pix in (("convnet + radial", spec2, True), ("image 16x7x7, pixel planes (usf_conv2d_pix)", spec3, True),
                         ("image 16x7x7, implicit-GEMM route (usf_conv2d_rows)", spec3, False)):
        p = O.random_params(sp, 9)
        xs = torch.rand(40, *sp["in_dims"], generator=g)
        image_engine.PIX_CONV = pix
        try:
            f = build_flow(sp, p, device="cuda:0", precision="fp32")
            e = rel_err(f.log_prob(xs.cuda()), O.flow_log_prob(xs, sp, p))
            f.sample([4])
        finally:
            image_engine.PIX_CONV = True
        print(f"[{tag}] log_prob err {e:.2e}", flush=True)
        assert e < 2e-5
    # usf_conv2d_pix on an image cut into row tiles (H*W > 256), 5 x 5 taps, and the narrow SIMT path
    from usflows_b200 import ops
    from usflows_b200.ops import Act
    a16 = torch.zeros(2 * 20 * 18, 64, dtype=torch.float16, device="cuda")
    ops.pix_encode(torch.rand(2 * 20 * 18, 24, generator=g).cuda(), 20 * 18, a16)
    w16 = image_engine._pix_weight((torch.randn(32, 25 * 24, generator=g) / 25).cuda(), 25, 24, 32, None)
    o32 = torch.empty(2 * 20 * 18, 32, device="cuda")
    ops.conv2d_pix(a16, 2, 20, 18, 5, 1, w16, torch.zeros(32, device="cuda"), 32, out_f32=o32, out16=torch.empty_like(a16))
    ops.linear(0, Act(5000, 16, f32=torch.rand(5000, 16, generator=g).cuda()), torch.rand(16, 16, generator=g).cuda(), None, 16, 16,
               out=Act(5000, 16, f32=torch.empty(5000, 16, device="cuda")))
    torch.cuda.synchronize()
    assert torch.isfinite(o32).all()
    # Chi-family radius (USF_NORM_GAMMA_MIXTURE_SQ): density and sampling
    spec_chi = dict(in_dims=[32], coupling_blocks=1, hidden_dims=[32], affine_conjugation=True, lu_transform=1, householder=0,
                    base="radial", p=2, norm="chi", df=32, chi_scale=1.5)
    pc = O.random_params(spec_chi, 3)
    xc = torch.rand(100, 32, generator=g)
    fc = build_flow(spec_chi, pc, device="cuda:0", precision="fp32")
    e = rel_err(fc.log_prob(xc.cuda()), O.flow_log_prob(xc, spec_chi, pc))
    fc.sample([50])
    print(f"[radial chi] log_prob err {e:.2e}", flush=True)
    assert e < 2e-5
    # fp64 tensor-core products of the weight preparation: 64 x 64 tiles with ragged edges (dense and both triangular
    # kinds), 128 x 128 tiles (grid >= 2 x SMs) on a short K
    for tri in (ops.TRI_NONE, ops.TRI_LOWER_UPPER, ops.TRI_UPPER_LOWER):
        a = torch.randn(200, 200, generator=g, dtype=torch.float64)
        b = torch.randn(200, 200, generator=g, dtype=torch.float64)
        if tri:
            a, b = (a.tril(), b.triu()) if tri == ops.TRI_LOWER_UPPER else (a.triu(), b.tril())
        o = torch.empty(200, 200, dtype=torch.float64, device="cuda")
        ops.matmul_f64(a.cuda(), b.cuda(), o, tri)
        assert rel_err(o, a @ b) < 1e-13
    a, b = torch.randn(2300, 20, generator=g, dtype=torch.float64), torch.randn(20, 2290, generator=g, dtype=torch.float64)
    o = torch.empty(2300, 2290, dtype=torch.float64, device="cuda")
    ops.matmul_f64(a.cuda(), b.cuda(), o)
    assert rel_err(o, a @ b) < 1e-13
    print("[matmul_f64] ok", flush=True)
    # training step (autograd contractions incl. transposes) on a small flow
    from usflows_b200 import training
    import usflows_b200 as U
    ft = build_flow(spec, params, device="cuda:0", precision="fp32")
    ts = training.TrainStep(ft, U.SophiaG(list(ft.parameters()), lr=1e-4, weight_decay=0.0), distributed=False)
    print("[train] loss", float(ts.step(x.cuda())), flush=True)
    torch.cuda.synchronize()
    print("sanitize pass OK", flush=True)


if __name__ == "__main__":
    main()
