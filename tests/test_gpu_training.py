"""Training step on the B200: loss / gradient parity of the tcgen05-backed autograd path against the CPU oracle, and
a short `Flow.fit` run.  (Data-parallel logic: tests/test_training_host.py over gloo; NCCL path: bench.py --train.)"""
import numpy as np
import pytest
import torch

from helpers import build_flow, load_case, rel_err
from oracle import flow_oracle as O

pytestmark = pytest.mark.gpu


def _oracle_grads(spec, params, x):
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in params.items()}
    loss = -O.flow_log_prob(x, spec, p).mean()
    loss.backward()
    return float(loss.detach()), {k: v.grad for k, v in p.items() if v.grad is not None}


@pytest.mark.parametrize("name", ["d6_hh_normal", "d100_h50_hh", "c2_d784"])
def test_gradients_match_the_oracle_on_device(name):
    from usflows_b200 import training
    spec, params, arr = load_case(name)
    spec = dict(spec, base="normal")          # smooth loss: see tests/test_training_host.py
    flow = build_flow(spec, params)
    x = arr["x"][:96]
    loss = -training.log_prob_autograd(flow, x.cuda()).mean()
    loss.backward()
    want_loss, want = _oracle_grads(spec, params, x)
    assert abs(float(loss) - want_loss) <= 2e-5 * max(1.0, abs(want_loss))
    got = dict(flow.named_parameters())
    checked = 0
    for key, g in want.items():
        if key not in got or got[key].grad is None:
            continue
        parts = key.split(".")
        ref = g
        if parts[0] == "trainable_layers" and parts[2] == "block_transform":
            cand = ".".join([parts[0], str(int(parts[1]) + 2), "transform"] + parts[2:])
            if cand in want and torch.equal(params[cand], params[key]):
                ref = ref + want[cand]
        if key.endswith("L_raw"):
            ref = ref.tril(-1)
        if key.endswith("U_raw"):
            ref = ref.triu()
        assert rel_err(got[key].grad, ref) <= 5e-4, key
        checked += 1
    assert checked >= 6


def test_fit_runs_and_learns_on_device():
    spec, params, arr = load_case("d100_h50_hh")
    flow = build_flow(spec, params)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2048, 100, generator=g)
    np.random.seed(0)
    l0 = float(-flow.log_prob(x.cuda()).mean())
    losses = flow.fit(torch.utils.data.TensorDataset(x), optim=torch.optim.Adam, optim_params=dict(lr=1e-3),
                      batch_size=256, epochs=3)
    assert losses[-1] < losses[0]
    assert float(-flow.log_prob(x.cuda()).mean()) < l0
    # the reference's default optimiser (sign momentum): every step moves every trained entry by lr
    flow2 = build_flow(spec, params)
    before = flow2.layers[-1].scale.detach().clone()
    flow2.fit(torch.utils.data.TensorDataset(x[:256]), optim_params=dict(lr=1e-4, weight_decay=0.0), batch_size=256)
    assert torch.allclose((flow2.layers[-1].scale.detach() - before).abs(), torch.full_like(before, 1e-4), atol=1e-7)
