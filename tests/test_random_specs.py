"""Seeded sweep over randomly drawn flow configurations (flat and image-shaped events, every conditioner and base family,
with / without conjugation, Householder, gating, LayerNorm, mask types): the launch planner on the emulated backend and the
exportable reference semantics against the CPU oracle.  Host logic only (no GPU); the oracle itself is pinned against the
real reference by tests/test_oracle.py."""
import random

import pytest
import torch

from helpers import build_flow, rel_err
from oracle import flow_oracle as O


def _draw(seed: int):
    rng = random.Random(seed)
    image = rng.random() < 0.4
    spec = dict(coupling_blocks=rng.randint(1, 3), affine_conjugation=rng.random() < 0.6, lu_transform=rng.randint(1, 2),
                householder=rng.choice([0, 0, 1, 2]), masktype=rng.choice(["checkerboard", "channel"]) if image else "checkerboard")
    if image:
        spec["in_dims"] = [rng.choice([4, 6, 8, 16]), rng.randint(3, 6), rng.randint(3, 6)]
        spec.update(conditioner="convnet2d", c_hidden=rng.choice([4, 8, 16]), num_layers=rng.randint(1, 3),
                    kernel_size=rng.choice([1, 3, 3, 5]), gating=rng.random() < 0.7, normalize_layers=rng.random() < 0.7)
    else:
        d = rng.choice([8, 12, 16, 24, 40])
        spec["in_dims"] = [d]
        if rng.random() < 0.5:
            widths = [rng.choice([8, 16, 24]) for _ in range(rng.randint(1, 3))]
            spec.update(conditioner="convnet", c_hidden=widths, gating=rng.random() < 0.7,
                        normalize_layers=rng.random() < 0.7)
        else:
            spec["hidden_dims"] = [rng.choice([8, 16, 32]) for _ in range(rng.randint(1, 3))]
    base = rng.choice(["laplace", "normal", "radial", "radial"])
    spec["base"] = base
    if base == "radial":
        spec.update(p=rng.choice([1, 2, "inf"]), norm=rng.choice(["lognormal", "gammamm"]), n_comp=rng.randint(2, 9))
    return spec


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_matches_oracle(fake_ops, seed):
    spec = _draw(seed)
    params = O.random_params(spec, 1000 + seed)
    g = torch.Generator().manual_seed(seed)
    n = 9
    x = torch.rand(n, *spec["in_dims"], generator=g)
    z0 = torch.randn(n, *spec["in_dims"], generator=g)
    want_lp = O.flow_log_prob(x, spec, params, dtype=torch.float64)
    want_z = O.flow_backward(x, spec, params, dtype=torch.float64)
    want_y = O.flow_forward(z0, spec, params, dtype=torch.float64)
    e_lp = rel_err(O.flow_log_prob(x, spec, params), want_lp)
    e_z = rel_err(O.flow_backward(x, spec, params), want_z)
    e_y = rel_err(O.flow_forward(z0, spec, params), want_y)
    for mode in ("fp32_simt", "fp32"):
        flow = build_flow(spec, params, device="cpu", precision=mode)
        assert rel_err(flow.log_prob(x), want_lp) <= 3 * e_lp + 2e-5, (spec, mode)
        assert rel_err(flow.backward(x), want_z) <= 3 * e_z + 5e-5, (spec, mode)
        assert rel_err(flow._forward(z0), want_y) <= 3 * e_y + 5e-5, (spec, mode)
        assert flow.sample([3]).shape == (3, *spec["in_dims"])
    module = flow.reference_module("log_prob")
    assert rel_err(module(x), want_lp) <= 3 * e_lp + 2e-5, spec
    assert rel_err(flow.reference_module("backward")(x), want_z) <= 3 * e_z + 5e-5, spec


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_matches_oracle_on_gpu(seed):
    """The same sweep through the C ABI on the device (default fp32 mode and the tf32-split engine)."""
    spec = _draw(seed)
    params = O.random_params(spec, 1000 + seed)
    g = torch.Generator().manual_seed(seed)
    n = 300
    x = torch.rand(n, *spec["in_dims"], generator=g)
    z0 = torch.randn(n, *spec["in_dims"], generator=g)
    want_lp = O.flow_log_prob(x, spec, params, dtype=torch.float64)
    want_z = O.flow_backward(x, spec, params, dtype=torch.float64)
    want_y = O.flow_forward(z0, spec, params, dtype=torch.float64)
    e_lp = rel_err(O.flow_log_prob(x, spec, params), want_lp)
    e_z = rel_err(O.flow_backward(x, spec, params), want_z)
    e_y = rel_err(O.flow_forward(z0, spec, params), want_y)
    for mode in ("fp32", "fp32_tf32"):
        flow = build_flow(spec, params, precision=mode)
        assert rel_err(flow.log_prob(x.cuda()), want_lp) <= 3 * e_lp + 1e-5, (spec, mode)
        assert rel_err(flow.backward(x.cuda()), want_z) <= 3 * e_z + 3e-5, (spec, mode)
        assert rel_err(flow._forward(z0.cuda()), want_y) <= 3 * e_y + 3e-5, (spec, mode)
    s = flow.sample([5])
    assert s.shape == (5, *spec["in_dims"]) and bool(torch.isfinite(s).all())


# ----------------------------------------------------------------------------------------------------------------------
# Second sweep: the families added late in round 2 -- networks.ConvNet's convolutional branch (varying widths, projected
# residuals), the context-conditioned conditioners of soft training (with and without an explicit context), simplify().
# ----------------------------------------------------------------------------------------------------------------------
def _draw_wide(seed: int):
    rng = random.Random(5000 + seed)
    image = rng.random() < 0.6
    soft = rng.random() < 0.5
    spec = dict(coupling_blocks=rng.randint(1, 3), affine_conjugation=rng.random() < 0.6, lu_transform=rng.randint(1, 2),
                householder=rng.choice([0, 0, 1, 2]), masktype=rng.choice(["checkerboard", "channel"]) if image else "checkerboard",
                soft_training=soft, gating=rng.random() < 0.7, normalize_layers=rng.random() < 0.7)
    if image:
        spec["in_dims"] = [rng.choice([4, 6, 8, 16]), rng.randint(3, 6), rng.randint(3, 6)]
        if rng.random() < 0.4:
            spec.update(conditioner="condconvnet2d" if soft else "convnet2d", c_hidden=rng.choice([4, 8, 16]),
                        num_layers=rng.randint(1, 3), kernel_size=rng.choice([1, 3, 3, 5]))
        else:
            spec.update(conditioner="condconvnet" if soft else "convnet", kernel_size=rng.choice([1, 3, 3]),
                        c_hidden=[rng.choice([4, 8, 12, 16]) for _ in range(rng.randint(1, 3))])
    else:
        spec["in_dims"] = [rng.choice([8, 12, 16, 24, 40])]
        spec.update(conditioner="condconvnet" if soft else "convnet",
                    c_hidden=[rng.choice([8, 16, 24]) for _ in range(rng.randint(1, 3))])
    base = rng.choice(["laplace", "normal", "radial"])
    spec["base"] = base
    if base == "radial":
        spec.update(p=rng.choice([1, 2, "inf"]), norm=rng.choice(["lognormal", "gammamm"]), n_comp=rng.randint(2, 9))
    return spec


def _check_wide(seed: int, device: str, n: int, modes, slack_lp: float, slack_z: float, draw=None):
    spec = (draw or _draw_wide)(seed)
    params = O.random_params(spec, 7000 + seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, *spec["in_dims"], generator=g)
    z0 = torch.randn(n, *spec["in_dims"], generator=g)
    ctx = torch.rand(n, 1, generator=g) * 2
    want_lp = O.flow_log_prob(x, spec, params, dtype=torch.float64)
    want_z = O.flow_backward(x, spec, params, dtype=torch.float64)
    want_y = O.flow_forward(z0, spec, params, dtype=torch.float64)
    e_lp = rel_err(O.flow_log_prob(x, spec, params), want_lp)
    e_z = rel_err(O.flow_backward(x, spec, params), want_z)
    e_y = rel_err(O.flow_forward(z0, spec, params), want_y)
    xd, zd = x.to(device), z0.to(device)
    for mode in modes:
        flow = build_flow(spec, params, device=device, precision=mode)
        assert rel_err(flow.log_prob(xd), want_lp) <= 3 * e_lp + slack_lp, (spec, mode)
        assert rel_err(flow.backward(xd), want_z) <= 3 * e_z + slack_z, (spec, mode)
        assert rel_err(flow._forward(zd), want_y) <= 3 * e_y + slack_z, (spec, mode)
        if spec["soft_training"]:
            cspec = dict(spec, _context=ctx.double())
            want_ctx = O.flow_log_prob(x, cspec, params, dtype=torch.float64)
            e_ctx = rel_err(O.flow_log_prob(x, dict(spec, _context=ctx), params), want_ctx)
            assert rel_err(flow.log_prob(xd, context=ctx.to(device)), want_ctx) <= 3 * e_ctx + slack_lp, (spec, mode)
    simple = flow.simplify()
    want_simple = want_lp
    if spec.get("conditioner") == "conddense" and spec["soft_training"]:
        # the reference's simplified flow is a plain `Flow` (flows.py:600-606): no zero context is substituted any more, so a
        # ConditionalDenseNN skips its context layer there
        want_simple = O.flow_log_prob(x, dict(spec, soft_training=False), params, dtype=torch.float64)
    assert rel_err(simple.log_prob(xd), want_simple) <= 6 * e_lp + 2 * slack_lp, spec
    return flow, spec, x, want_lp, e_lp


@pytest.mark.parametrize("seed", range(16))
def test_random_wide_configuration_matches_oracle(fake_ops, seed):
    flow, spec, x, want_lp, e_lp = _check_wide(seed, "cpu", 9, ("fp32_simt", "fp32"), 2e-5, 5e-5)
    assert rel_err(flow.reference_module("log_prob")(x), want_lp) <= 3 * e_lp + 2e-5, spec


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(16))
def test_random_wide_configuration_matches_oracle_on_gpu(seed):
    flow, spec, _, _, _ = _check_wide(seed, "cuda", 300, ("fp32", "fp32_tf32"), 1e-5, 3e-5)
    s = flow.sample([5])
    assert s.shape == (5, *spec["in_dims"]) and bool(torch.isfinite(s).all())


# ----------------------------------------------------------------------------------------------------------------------
# Third sweep: what the end of round 2 added -- networks.ConditionalDenseNN (inside and outside soft training: zero context
# in log_prob, none in backward / _forward), every radius distribution of the Lp-radial base (the reference's Chi / Gamma /
# WeibullMM / LogNormalMM, torch's Chi2 / HalfNormal / Weibull / Exponential / LogNormal objects).
# ----------------------------------------------------------------------------------------------------------------------
_NORMS = ["lognormal", "gammamm", "gamma", "chi", "chi2", "halfnormal", "weibull", "exponential", "torchlognormal",
          "weibullmm", "lognormalmm"]


def _draw_late(seed: int):
    rng = random.Random(9000 + seed)
    image = rng.random() < 0.3
    spec = dict(coupling_blocks=rng.randint(1, 3), affine_conjugation=rng.random() < 0.6, lu_transform=rng.randint(1, 2),
                householder=rng.choice([0, 0, 1, 2]), masktype=rng.choice(["checkerboard", "channel"]) if image else "checkerboard",
                soft_training=False)
    if image:
        spec["in_dims"] = [rng.choice([4, 6, 8]), rng.randint(3, 5), rng.randint(3, 5)]
        spec.update(conditioner="convnet2d", c_hidden=rng.choice([4, 8]), num_layers=rng.randint(1, 2),
                    kernel_size=rng.choice([1, 3]), gating=rng.random() < 0.7, normalize_layers=rng.random() < 0.7)
    else:
        spec["in_dims"] = [rng.choice([8, 12, 16, 24, 40])]
        spec["hidden_dims"] = [rng.choice([8, 16, 32]) for _ in range(rng.randint(1, 3))]
        if rng.random() < 0.7:
            spec.update(conditioner="conddense", soft_training=rng.random() < 0.6)
    d = 1
    for n in spec["in_dims"]:
        d *= n
    spec.update(base="radial", p=rng.choice([1, 2, "inf"]), norm=_NORMS[seed % len(_NORMS)], n_comp=rng.randint(2, 7),
                df=rng.choice([d, d / 2 + 0.5, 3.0]), chi_scale=rng.choice([0.5, 1.0, 2.5]), w_scale=0.5 * d ** 0.5 + 3 * rng.random(),
                w_conc=1.2 + 2 * rng.random(), rate=0.1 + rng.random(), ln_loc=0.3 + rng.random(), ln_scale=0.2 + 0.5 * rng.random())
    if rng.random() < 0.25:
        spec["base"] = rng.choice(["laplace", "normal"])
    return spec


@pytest.mark.parametrize("seed", range(22))
def test_random_late_configuration_matches_oracle(fake_ops, seed):
    flow, spec, x, want_lp, e_lp = _check_wide(seed, "cpu", 9, ("fp32_simt", "fp32"), 2e-5, 5e-5, draw=_draw_late)
    assert rel_err(flow.reference_module("log_prob")(x), want_lp) <= 3 * e_lp + 2e-5, spec


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(22))
def test_random_late_configuration_matches_oracle_on_gpu(seed):
    flow, spec, _, _, _ = _check_wide(seed, "cuda", 300, ("fp32", "fp32_tf32"), 1e-5, 3e-5, draw=_draw_late)
    s = flow.sample([5])
    assert s.shape == (5, *spec["in_dims"]) and bool(torch.isfinite(s).all())
