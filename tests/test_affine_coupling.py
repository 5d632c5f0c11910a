"""Affine (scale-and-shift) coupling -- an EXTENSION the task names; the reference's coupling is additive only, so no
reference output exists for it ("parity unpinned").  Checked here: (1) the CPU restatement (oracle) against first
principles -- round trip and the autograd Jacobian in fp64; (2) the product's host logic (planner, mask compression,
per-row log-det bookkeeping) against that restatement on the emulated backend; GPU parity is in test_gpu_parity.py."""
import pytest
import torch

from helpers import build_flow, rel_err
from oracle import flow_oracle as O

SPECS = {
    "d8_h32": dict(in_dims=[8], coupling_blocks=2, hidden_dims=[32, 32], affine_conjugation=True, lu_transform=1,
                   householder=0, base="laplace", coupling="affine"),
    "d16_h48": dict(in_dims=[16], coupling_blocks=2, hidden_dims=[48, 48], affine_conjugation=True, lu_transform=1,
                    householder=0, base="normal", coupling="affine"),
    "d6_noconj": dict(in_dims=[6], coupling_blocks=2, hidden_dims=[16], affine_conjugation=False, lu_transform=1,
                      householder=0, base="laplace", coupling="affine"),
}


def _case(name, rows=40):
    spec = SPECS[name]
    params = O.random_params(spec, 11)
    g = torch.Generator().manual_seed(5)
    return spec, params, torch.rand(rows, spec["in_dims"][0], generator=g)


@pytest.mark.parametrize("name", sorted(SPECS))
def test_oracle_round_trip_and_jacobian(name):
    spec, params, x = _case(name, rows=6)
    z = O.flow_backward(x, spec, params, torch.float64)
    assert rel_err(O.flow_forward(z, spec, params, torch.float64), x) < 1e-6      # fp64; the random LU layers are ill-conditioned
    # log p(x) = base(z) + log|det dz/dx|: the second term from the autograd Jacobian of the data -> latent map
    lp = O.flow_log_prob(x, spec, params, torch.float64)
    p64 = O._cast(params, torch.float64)
    for i in range(x.shape[0]):
        J = torch.autograd.functional.jacobian(lambda v: O.flow_backward(v[None], spec, params, torch.float64)[0], x[i].double())
        want = O.base_log_prob(z[i:i + 1], spec, p64)[0] + torch.linalg.slogdet(J)[1]
        assert abs(float(lp[i] - want)) < 1e-6 * max(1.0, abs(float(want)))


@pytest.mark.parametrize("mode", ["fp32_simt", "fp32", "fp32_tf32"])
@pytest.mark.parametrize("name", sorted(SPECS))
def test_host_logic_matches_the_restatement(fake_ops, name, mode):
    import fake_backend
    spec, params, x = _case(name)
    flow = build_flow(spec, params, device="cpu", precision=mode)
    fake_backend.CALLS.clear()
    lp = flow.log_prob(x)
    assert any(c[0] == "affine_couple" for c in fake_backend.CALLS)
    assert rel_err(lp, O.flow_log_prob(x, spec, params, torch.float64)) < 3e-5
    z = flow.backward(x)
    assert rel_err(z, O.flow_backward(x, spec, params, torch.float64)) < 5e-5
    z0 = torch.randn(x.shape, generator=torch.Generator().manual_seed(9))      # (a round trip in fp32 only measures the
    assert rel_err(flow._forward(z0), O.flow_forward(z0, spec, params, torch.float64)) < 5e-5   # conditioning of the random LU layers)
    assert torch.equal(x, _case(name)[2])                       # the caller's tensor is never updated in place


def test_layer_log_det_and_reference_module():
    spec, params, x = _case("d8_h32")
    flow = build_flow(spec, params, device="cpu")
    lp_ref = flow.reference_module("log_prob")(x)
    assert rel_err(lp_ref, O.flow_log_prob(x, spec, params, torch.float64)) < 2e-5
    assert rel_err(flow.reference_module("forward")(x), O.flow_forward(x, spec, params, torch.float64)) < 2e-5
    assert rel_err(flow.reference_module("backward")(x), O.flow_backward(x, spec, params, torch.float64)) < 2e-5


def test_constructor_checks():
    import usflows_b200 as U
    with pytest.raises(ValueError, match="2\\*d"):
        U.MaskedAffineCoupling(torch.tensor([1.0, 0.0, 1.0, 0.0]), U.DenseNN(4, [8], param_dims=[4]))
    with pytest.raises(ValueError, match="coupling"):
        build_flow(dict(SPECS["d8_h32"], coupling="multiplicative"), {}, device="cpu")
