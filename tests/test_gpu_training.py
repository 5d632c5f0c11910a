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


ENGINE_SPECS = {
    "c2_shape": dict(in_dims=[784], coupling_blocks=2, hidden_dims=[1024, 1024], affine_conjugation=True, lu_transform=1,
                     householder=0, base="laplace"),
    "d64_normal": dict(in_dims=[64], coupling_blocks=2, hidden_dims=[64, 48], affine_conjugation=True, lu_transform=1,
                       householder=0, base="normal"),
    "d96_3layer": dict(in_dims=[96], coupling_blocks=3, hidden_dims=[40, 56, 32], affine_conjugation=True,
                       lu_transform=1, householder=0, base="laplace"),
    "d48_noconj": dict(in_dims=[48], coupling_blocks=1, hidden_dims=[64, 32], affine_conjugation=False,
                       lu_transform=1, householder=0, base="laplace"),
}


def _oracle_grads64(spec, params, x):
    p = {k: (v.double() if v.is_floating_point() else v).clone().requires_grad_(v.is_floating_point()) for k, v in params.items()}
    loss = -O.flow_log_prob(x.double(), spec, p, dtype=torch.float64).mean()
    loss.backward()
    return float(loss.detach()), {k: v.grad for k, v in p.items() if v.grad is not None}


def _folded(want, params, key):
    """Reference gradient of a product parameter: contributions of the aliased InverseTransform copy added, LU masks applied."""
    parts = key.split(".")
    ref = want[key]
    if parts[0] == "trainable_layers" and parts[2] == "block_transform":
        cand = ".".join([parts[0], str(int(parts[1]) + 2), "transform"] + parts[2:])
        if cand in want and torch.equal(params[cand], params[key]):
            ref = ref + want[cand]
    if key.endswith("L_raw"):
        ref = ref.tril(-1)
    if key.endswith("U_raw"):
        ref = ref.triu()
    return ref


@pytest.mark.parametrize("name,rows", [("d64_normal", 96), ("d96_3layer", 300), ("d48_noconj", 200), ("c2_shape", 512)])
def test_hand_written_training_pass_matches_the_oracle_on_device(name, rows):
    """train_engine.TrainEngine on the B200 (fp16-split tcgen05 contractions forward / dX / split-K dW, batched triangular
    inverses, glue kernels) against autograd through the CPU oracle."""
    from usflows_b200 import train_engine
    spec = ENGINE_SPECS[name]
    params = O.random_params(spec, 21)
    flow = build_flow(spec, params)
    assert train_engine.supports(flow)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(rows, spec["in_dims"][0], generator=g)
    eng = train_engine.TrainEngine(flow, rows)
    for it in range(2):                                   # the second pass re-uses every buffer
        loss = eng.step(x.cuda(), rows)
        assert int(eng.flag) == 0
    want_loss, want = _oracle_grads(spec, params, x)
    assert abs(float(loss) - want_loss) <= 2e-5 * max(1.0, abs(want_loss))
    # yardstick: the fp64 evaluation of the same loss.  The reference's own fp32 autograd sits up to ~1e-3 from it on the
    # LU gradients of these random stacks (W^-1 enters twice, conditioning ~1e2 per layer), so the bound is the larger of
    # 5e-4 and 3x the reference's own fp32 error, per parameter.
    _, truth = _oracle_grads64(spec, params, x)
    got = dict(flow.named_parameters())
    checked = 0
    for key in want:
        if key not in got or got[key].grad is None:
            continue
        t64 = _folded(truth, params, key)
        ref_err = rel_err(_folded(want, params, key), t64)
        err = rel_err(got[key].grad, t64)
        assert err <= max(5e-4, 3 * ref_err), (key, err, ref_err)
        checked += 1
    assert checked >= 6


def test_train_step_engine_and_autograd_routes_agree_and_fit_learns():
    import usflows_b200 as U
    from usflows_b200 import training
    spec = ENGINE_SPECS["d64_normal"]
    params = O.random_params(spec, 3)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(256, 64, generator=g).cuda()
    grads = []
    for engine in (True, False):
        flow = build_flow(spec, params)
        ts = training.TrainStep(flow, torch.optim.SGD(flow.parameters(), lr=0.0), distributed=False, engine=engine)
        loss = ts.step(x)
        grads.append((float(loss), {k: p.grad.clone() for k, p in flow.named_parameters()}))
        assert float(ts.infeasible) == 0 and float(ts.out_of_range) == 0
    assert abs(grads[0][0] - grads[1][0]) <= 1e-5 * max(1.0, abs(grads[1][0]))
    for k in grads[0][1]:
        assert rel_err(grads[0][1][k], grads[1][1][k]) <= 5e-4, k
    flow = build_flow(spec, params)
    np.random.seed(0)
    xs = torch.rand(2048, 64, generator=g)
    l0 = float(-flow.log_prob(xs.cuda()).mean())
    losses = flow.fit(torch.utils.data.TensorDataset(xs), optim=torch.optim.Adam, optim_params=dict(lr=1e-3), batch_size=256,
                      epochs=3)
    assert losses[-1] < losses[0] and float(-flow.log_prob(xs.cuda()).mean()) < l0


def test_captured_training_step_matches_the_eager_one():
    """TrainStep replays the whole step (hand-written pass, SophiaG, invertibility counter) as one CUDA graph after two
    eager steps: same loss trajectory as the eager route (split-K reductions are atomic, so gradients agree to rounding and
    the sign-momentum update may differ by one lr on entries whose momentum is at rounding distance from zero)."""
    import usflows_b200 as U
    from usflows_b200 import training
    spec = ENGINE_SPECS["d64_normal"]
    params = O.random_params(spec, 3)
    g = torch.Generator().manual_seed(1)
    xs = [torch.rand(256, 64, generator=g).cuda() for _ in range(6)]
    runs = []
    for graph in (True, False):
        flow = build_flow(spec, params)
        ts = training.TrainStep(flow, U.SophiaG(list(flow.parameters()), lr=1e-4, weight_decay=0.0), distributed=False,
                                graph=graph)
        losses = [float(ts.step(x)) for x in xs]
        assert float(ts.infeasible) == 0 and float(ts.out_of_range) == 0
        assert (len(ts._graphs) == 1) is graph
        runs.append((losses, {k: v.detach().clone() for k, v in flow.state_dict().items()}))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b))
    assert runs[0][0][-1] < runs[0][0][0]                      # it learns
    for k, v in runs[1][1].items():
        assert float((runs[0][1][k] - v).abs().max()) <= 2 * 6 * 1e-4 + 1e-7, k
