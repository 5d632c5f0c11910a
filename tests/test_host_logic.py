"""Host-side logic (planning, fusion, plane bookkeeping, weight-prep order, caching, state-dict layout) tested on
CPU against the golden fixtures, with the C ABI replaced by a torch emulation (tests/fake_backend.py)."""
import pytest
import torch

import fake_backend
from helpers import (EXT_CASES, IMG_CASES, LARGE_CASES, SIMPLIFY_CASES, SMALL_CASES, SOFT_CASES, build_flow, layer_kinds,
                     load_case, load_simplify_case, rel_err)


@pytest.mark.parametrize("mode", ["fp32_simt", "fp32", "fp32_tf32", "tf32"])
@pytest.mark.parametrize("name", SMALL_CASES + ["c2_d784"] + EXT_CASES + IMG_CASES + SOFT_CASES)
def test_flow_program_matches_reference(fake_ops, name, mode):
    spec, params, arr = load_case(name)
    flow = build_flow(spec, params, device="cpu", precision=mode)
    lp = flow.log_prob(arr["x"])
    z = flow.backward(arr["x"])
    y = flow._forward(arr["z0"])
    tol = 2e-5 if mode != "tf32" else 1e-4       # the emulated tf32 engine multiplies in fp32
    assert rel_err(lp, arr["lp32"]) < tol
    assert rel_err(z, arr["z32"]) < 5e-5
    assert rel_err(y, arr["y32"]) < 5e-5


def test_bf16_mode_runs_with_its_own_bound(fake_ops):
    spec, params, arr = load_case("d32_h64")
    flow = build_flow(spec, params, device="cpu", precision="bf16")
    assert rel_err(flow.log_prob(arr["x"]), arr["lp32"]) < 5e-2


def test_kernel_count_per_log_prob(fake_ops):
    """one ingest + (2B+1) affine contractions (Scale folded into the first) + 3B conditioner contractions + one
    base-density reduction for a conjugated USFlow with B blocks; with `MERGE_AFFINE` the Aff_i^-1 . Aff_{i+1}
    pairs become one operator each (B+1 affine contractions)."""
    spec, params, arr = load_case("d100_h50_hh")
    flow = build_flow(spec, params, device="cpu", precision="fp32")
    flow.log_prob(arr["x"])               # includes weight preparation
    fake_backend.CALLS.clear()
    flow.log_prob(arr["x"])               # steady state: prepared weights are cached
    kinds = [c[0] for c in fake_backend.CALLS]
    B = spec["coupling_blocks"]
    assert spec["affine_conjugation"] and len(spec["hidden_dims"]) == 2
    assert kinds.count("ingest") == 1 and kinds.count("base_logprob") == 1
    assert kinds.count("linear") == (2 * B + 1) + 3 * B
    from usflows_b200 import engine
    engine.MERGE_AFFINE = True
    try:
        merged = build_flow(spec, params, device="cpu", precision="fp32")
        lp = merged.log_prob(arr["x"])
        fake_backend.CALLS.clear()
        merged.log_prob(arr["x"])
        assert [c[0] for c in fake_backend.CALLS].count("linear") == (B + 1) + 3 * B
        assert rel_err(lp, arr["lp32"]) < 5e-5
    finally:
        engine.MERGE_AFFINE = False


def test_mask_compression_halves_the_outer_conditioner_contractions(fake_ops):
    """d = 784: the checkerboard partition gives two 392-wide contiguous segments; the first conditioner GEMM
    reads K = 392 columns and the last writes N = 392 columns in place."""
    spec, params, arr = load_case("c2_d784")
    flow = build_flow(spec, params, device="cpu", precision="fp32")
    flow.log_prob(arr["x"][:4])
    fake_backend.CALLS.clear()
    flow.log_prob(arr["x"][:4])
    shapes = [(c[3], c[4]) for c in fake_backend.CALLS if c[0] == "linear"]      # (N, K)
    H = spec["hidden_dims"][0]
    assert shapes.count((784, 784)) == 2 * spec["coupling_blocks"] + 1
    assert shapes.count((H, 392)) == spec["coupling_blocks"] and shapes.count((392, H)) == spec["coupling_blocks"]


def test_fp16_range_guard_falls_back_to_the_tf32_split(fake_ops):
    """Inputs that push activations past the fp16 range raise the device flag; the chunk is recomputed with the
    tf32-split engine, so the result equals the "fp32_tf32" mode bit for bit."""
    from usflows_b200.ops import ENGINE_TC_3XF16, ENGINE_TC_3XTF32
    spec, params, arr = load_case("d100_h50_hh")
    x = arr["x"] * 3.0e5
    f16 = build_flow(spec, params, device="cpu", precision="fp32")
    tf = build_flow(spec, params, device="cpu", precision="fp32_tf32")
    fake_backend.CALLS.clear()
    lp = f16.log_prob(x)
    engines = {c[1] for c in fake_backend.CALLS if c[0] == "linear"}
    assert engines == {ENGINE_TC_3XF16, ENGINE_TC_3XTF32}
    assert torch.equal(lp, tf.log_prob(x))
    # in-range inputs never touch the fallback
    fake_backend.CALLS.clear()
    f16.log_prob(arr["x"])
    assert {c[1] for c in fake_backend.CALLS if c[0] == "linear"} == {ENGINE_TC_3XF16}


def test_prepared_weights_follow_weight_version(fake_ops):
    spec, params, arr = load_case("d6_hh_normal")
    flow = build_flow(spec, params, device="cpu", precision="fp32_simt")
    lp0 = flow.log_prob(arr["x"])
    with torch.no_grad():
        flow.layers[-1].scale.mul_(2.0)
    lp1 = flow.log_prob(arr["x"])
    assert not torch.allclose(lp0, lp1)


def test_state_dict_layout_matches_reference():
    spec, params, _ = load_case("d6_hh_normal")
    flow = build_flow(spec, params, device="cpu")
    assert set(flow.state_dict().keys()) == set(params.keys())
    for k, v in flow.state_dict().items():
        assert v.shape == params[k].shape, k
    # InverseTransform aliases block parameters (same storage), as in the reference
    sd = flow.state_dict()
    a = sd["trainable_layers.0.block_transform.transforms.0.L_raw"]
    b = sd["trainable_layers.2.transform.block_transform.transforms.0.L_raw"]
    assert a.data_ptr() == b.data_ptr()


def test_layer_stack_and_masks():
    import usflows_b200 as U
    spec, params, _ = load_case("d5_noconj")
    flow = build_flow(spec, params, device="cpu")
    names = [type(l).__name__ for l in flow.layers]
    assert names == ["BlockAffineTransform", "MaskedCoupling"] * 3 + ["BlockAffineTransform", "ScaleTransform"]
    assert torch.equal(U.USFlow.create_checkerboard_mask([4]), torch.tensor([[0., 1., 0., 1.]]))
    assert torch.equal(U.USFlow.create_checkerboard_mask([2, 2]), torch.tensor([[[0., 1.], [1., 0.]]]))
    assert torch.equal(U.USFlow.create_channel_mask([2, 2]), torch.tensor([[[0., 0.], [1., 1.]]]))


def test_cpu_tensor_is_rejected_without_fake_backend():
    spec, params, arr = load_case("d5_noconj")
    flow = build_flow(spec, params, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        flow.log_prob(arr["x"])


def test_individual_layers(fake_ops):
    """known-answer tests of the reference (tests/veriflow/transforms_test.py:5-67) through the engine."""
    import usflows_b200 as U
    dim = 10
    t = U.ScaleTransform([dim])
    with torch.no_grad():
        t.scale.copy_(torch.ones(dim) * 2)
    x = torch.ones(1, dim)
    y = t(x)
    assert (y == 2 * x).all() and (t.backward(y) == x).all()
    assert abs(float(t.log_abs_det_jacobian(x, y)) - dim * float(torch.log(torch.tensor(2.0)))) < 1e-6

    t = U.LUTransform(dim)
    with torch.no_grad():
        t.L_raw.copy_(torch.tril(torch.ones(dim, dim)))
        t.U_raw.copy_(torch.eye(dim))
        t.bias_vector.copy_(torch.zeros(dim))
    y = t(x)
    assert (y == (torch.arange(dim) + 1.0)).all() and (t.backward(y) == x).all()
    assert float(t.log_abs_det_jacobian(x, y)) == 0

    t = U.LeakyReLUTransform()
    x = torch.tensor([[1.0, -1.0] * 5])
    y = t(x)
    assert (y == x * torch.tensor([1.0, 0.01] * 5)).all()
    assert torch.allclose(t.backward(y), x)
    assert abs(float(t.log_abs_det_jacobian(x, y)) - 5 * float(torch.log(torch.tensor(0.01)))) < 1e-5

    t = U.Permute(torch.arange(dim))
    x = torch.arange(dim, dtype=torch.float32).reshape(1, dim)
    assert (t(x) == x).all() and (t.backward(x) == x).all()
    assert float(t.log_abs_det_jacobian(x, x)) == 0


@pytest.mark.parametrize("name,fused", [("c1_d2_laplace", True), ("d6_hh_normal", True), ("d5_noconj", True),
                                         ("d32_h64", False)])
def test_tiny_flows_run_as_one_launch(fake_ops, name, fused):
    """d <= 8 with conditioner width <= 64: the whole layer stack is one usf_flow_small launch (+ the base density);
    wider flows keep the per-layer contractions.  Switching the fusion off gives the same numbers."""
    import fake_backend
    from usflows_b200 import engine
    spec, params, arr = load_case(name)
    flow = build_flow(spec, params, device="cpu", precision="fp32")
    fake_backend.CALLS.clear()
    lp = flow.log_prob(arr["x"])
    kinds = [c[0] for c in fake_backend.CALLS]
    assert ("flow_small" in kinds) == fused
    if fused:
        assert kinds.count("flow_small") == 1 and "linear" not in kinds and kinds.count("base_logprob") == 1
        engine.FUSE_SMALL = False
        try:
            flow2 = build_flow(spec, params, device="cpu", precision="fp32")
            assert rel_err(flow2.log_prob(arr["x"]), lp) < 2e-5      # fp32 both ways, different summation order
            assert rel_err(flow2._forward(arr["z0"]), flow._forward(arr["z0"])) < 2e-5
        finally:
            engine.FUSE_SMALL = True


def test_replicas_of_the_row_shard_driver_are_independent_and_keep_the_aliasing(fake_ops):
    """parallel.replicate: a replica shares no storage with the source flow, keeps the InverseTransform <-> block parameter
    aliasing (flows.py:469-470) and evaluates to the same bits; row blocks cover every row once."""
    from usflows_b200 import parallel
    spec, params, arr = load_case("d100_h50_hh")
    flow = build_flow(spec, params, device="cpu", precision="fp32")
    want = flow.log_prob(arr["x"])                                   # fills the caches a replica must not copy
    rep = parallel.replicate(flow, "cpu")
    assert torch.equal(rep.log_prob(arr["x"]), want)
    src = {id(p): n for n, p in flow.named_parameters()}
    for n, p in rep.named_parameters():
        assert id(p) not in src and p.data_ptr() != dict(flow.named_parameters())[n].data_ptr()
    inv = [l for l in rep.layers if type(l).__name__ == "InverseTransform"]
    assert inv and all(any(l.transform is b for b in rep.layers) for l in inv)          # aliasing preserved
    with torch.no_grad():
        next(rep.parameters()).add_(0.25)
    assert torch.equal(flow.log_prob(arr["x"]), want)                 # the source is untouched
    for n, world in ((10, 3), (7, 8), (0, 2), (65536, 8)):
        cuts = [parallel.shard_bounds(n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))


@pytest.mark.parametrize("name", SIMPLIFY_CASES)
def test_simplify_matches_the_reference_simplify(fake_ops, name):
    """`Flow.simplify()` (flows.py:600-606): same layer classes and state-dict keys as the reference's simplified flow,
    same log_prob / latents / samples (golden outputs of the reference's simplified flow), and idempotent."""
    spec, params, arr = load_case(name)
    meta, want = load_simplify_case(name)
    flow = build_flow(spec, params, device="cpu", precision="fp32")
    simple = flow.simplify()
    assert layer_kinds(simple) == meta["layers"]
    assert list(simple.state_dict().keys()) == meta["state_keys"]
    assert rel_err(simple.log_prob(arr["x"]), want["lp"]) < 2e-5
    assert rel_err(simple.backward(arr["x"]), want["z"]) < 5e-5
    assert rel_err(simple._forward(arr["z0"]), want["y"]) < 5e-5
    again = simple.simplify()
    assert layer_kinds(again) == meta["layers"]
    assert torch.equal(again.log_prob(arr["x"]), simple.log_prob(arr["x"]))
    assert simple.is_feasible()
    assert rel_err(simple.reference_module("log_prob")(arr["x"]), want["lp"]) < 2e-5     # export of a simplified flow


def test_plane_linear_and_1x1_conv_built_directly(fake_ops):
    """`PlaneBijectiveLinearTransform(dim, m, bias, m_inv)` (transforms.py:618-695) and `Bijective1x1Conv2d(weight, bias)`
    (transforms.py:1031-1176) from plain tensors: forward / backward invert each other, log|det| as the reference's."""
    import usflows_b200 as U
    g = torch.Generator().manual_seed(5)
    d = 6
    m = torch.randn(d, d, generator=g) + 3 * torch.eye(d)
    b = torch.randn(d, generator=g)
    t = U.PlaneBijectiveLinearTransform(d, m, b, torch.linalg.inv(m))
    x = torch.randn(9, d, generator=g)
    y = t.forward(x)
    assert rel_err(y, x @ m.t() + b) < 1e-6
    assert rel_err(t.backward(y), x) < 1e-5
    assert abs(float(t.log_abs_det_jacobian(x, y)) - float(torch.linalg.slogdet(m)[1])) < 1e-5
    assert list(t.state_dict().keys()) == ["forth.weight", "forth.bias", "back.weight", "back.bias"]
    assert torch.equal(t.matrix(), m) and torch.equal(t.bias(), b)
    t2 = U.PlaneBijectiveLinearTransform(d, m, b)                # inverse derived here (the reference requires it)
    assert rel_err(t2.inverse_matrix(), torch.linalg.inv(m)) < 1e-6

    C, H, W = 4, 3, 5
    w = torch.randn(C, C, generator=g) + 2 * torch.eye(C)
    cb = torch.randn(C, generator=g)
    conv = U.Bijective1x1Conv2d(w.view(C, C, 1, 1), cb)
    xi = torch.randn(7, C, H, W, generator=g)
    ladj = conv.log_abs_det_jacobian(xi, xi)
    assert ladj.shape == (7,) and rel_err(ladj, torch.full((7,), float(torch.linalg.slogdet(w)[1]) * H * W)) < 1e-6
    flow = U.Flow(U.Normal(torch.zeros(C, H, W), torch.ones(C, H, W)), [conv], device="cpu")
    assert conv.n_blocks == H * W                                # bound from the flow's event shape
    z = flow.backward(xi)
    want = torch.nn.functional.conv2d(xi - cb.view(1, C, 1, 1), torch.linalg.inv(w).view(C, C, 1, 1))
    assert rel_err(z, want) < 1e-5
    assert rel_err(flow._forward(z), xi) < 1e-5
    base = torch.distributions.Normal(0.0, 1.0).log_prob(want).sum((1, 2, 3))
    assert rel_err(flow.log_prob(xi), base - ladj) < 1e-5
    with pytest.raises(ValueError):
        U.Bijective1x1Conv2d(torch.zeros(3, 4, 1, 1))


@pytest.mark.parametrize("name", SOFT_CASES)
def test_soft_training_context_matches_the_reference(fake_ops, name):
    """USFlow(soft_training=True) over CondConvNet2D / CondConvNet conditioners (flows.py:559-565, networks.py:513-680):
    `log_prob(x)` is the context-0 evaluation (covered by test_flow_program_matches_reference through the kernels'
    launch program), `log_prob(x, context)` reproduces the reference's output for per-sample contexts."""
    spec, params, arr = load_case(name)
    flow = build_flow(spec, params, device="cpu", precision="fp32")
    assert flow.soft_training and float(flow.training_noise_prior.high) == pytest.approx(0.01)
    lp_ctx = flow.log_prob(arr["x"], context=arr["ctx"])
    assert rel_err(lp_ctx, arr["lp32_ctx"]) < 2e-5
    assert rel_err(lp_ctx, arr["lp32"]) > 1e-7                       # the context does reach the conditioners
    zero = flow.log_prob(arr["x"], context=torch.zeros(arr["x"].shape[0], 1))
    assert rel_err(zero, flow.log_prob(arr["x"])) < 2e-6              # explicit zeros = the implicit default
    y0 = flow.sample([4], context=torch.zeros(4, 1))
    assert y0.shape == (4, *spec["in_dims"]) and bool(torch.isfinite(y0).all())


def test_soft_training_needs_a_conditional_conditioner(fake_ops):
    """The reference hands the context to the conditioner, which pyro's DenseNN rejects with a TypeError."""
    spec, params, arr = load_case("d32_h64")
    flow = build_flow(dict(spec, soft_training=True), params, device="cpu")
    with pytest.raises(TypeError):
        flow.log_prob(arr["x"])
    plain = build_flow(spec, params, device="cpu")                    # USFlow drops a context unless soft_training
    assert torch.equal(plain.log_prob(arr["x"], context=torch.ones(arr["x"].shape[0], 1)), plain.log_prob(arr["x"]))


def test_rotation_and_block_lu_layers_match_the_reference(fake_ops):
    """`Rotation`, `CompositeRotation` (transforms.py:476-616) and `BlockLUTransform` (transforms.py:1488-1622) against
    outputs of the reference (tests/golden/layers.npz)."""
    from helpers import check_standalone_layers
    check_standalone_layers("cpu")


@pytest.mark.parametrize("name", ["d32_h64", "d6_hh_normal"])
def test_plain_torch_distribution_object_as_the_base(fake_ops, name):
    """`Flow` takes a plain torch Laplace / Normal object (the reference's older configurations pass
    `pyro.distributions.Laplace(loc, scale)`; flows.py:97-101 turns its batch dims into event dims): same log-probs as the
    module base, no parameters, no state-dict keys of the base."""
    import usflows_b200 as U
    spec, params, arr = load_case(name)
    flow = build_flow(spec, params, device="cpu", precision="fp32")
    b = flow.base_distribution.base_dist if hasattr(flow.base_distribution, "base_dist") else flow.base_distribution
    scale = torch.nn.functional.softplus(b.scale_unconstrained.detach())
    cls = torch.distributions.Laplace if spec["base"] == "laplace" else torch.distributions.Normal
    for dist in (cls(b.loc.detach(), scale), torch.distributions.Independent(cls(b.loc.detach(), scale), 1),
                 U.Independent(cls(b.loc.detach(), scale), 1)):          # the reference's own wrapper (distributions.py:709-728)
        plain = U.Flow(dist, flow.layers, device="cpu", precision="fp32")
        assert rel_err(plain.log_prob(arr["x"]), arr["lp32"]) < 2e-5
        assert not any(k.startswith("base_distribution") for k in plain.state_dict())
        assert not any(n.startswith("base_distribution") for n, _ in plain.named_parameters())
        assert tuple(plain.sample([7]).shape) == (7, *spec["in_dims"])
    with pytest.raises(NotImplementedError):
        U.Flow(torch.distributions.Cauchy(torch.zeros(4), torch.ones(4)), flow.layers)


def test_conditional_dense_nn_contract(fake_ops):
    """`ConditionalDenseNN` (networks.py:681-752): reference layer list / state-dict keys; no context skips the context
    layer, a context adds `L1 c` before the first ReLU; a soft-training USFlow lowers `log_prob` with the zero context
    (the context layer's bias stays) and `backward` without one -- two launch programs per direction."""
    import usflows_b200 as U
    torch.manual_seed(0)
    net = U.ConditionalDenseNN(12, 1, [16, 8], 12)
    assert list(net.state_dict()) == [f"layers.{i}.{n}" for i in range(4) for n in ("weight", "bias")]
    assert net.layers[1].weight.shape == (16, 1) and net.context_channels == 1
    x, c = torch.randn(9, 12), torch.rand(9, 1)
    L = net.layers
    def ref(ctx):
        h = L[0](x) if ctx is None else L[0](x) + L[1](ctx)
        return L[3](torch.relu(L[2](torch.relu(h))))
    assert rel_err(net(x), ref(None).detach()) < 1e-6
    assert rel_err(net(x, c), ref(c).detach()) < 1e-6
    assert rel_err(net(x, torch.zeros(9, 1)), ref(torch.zeros(9, 1)).detach()) < 1e-6
    with pytest.raises(NotImplementedError):
        U.ConditionalDenseNN(12, 1, [16], 12, nonlinearity=torch.nn.Tanh())
    spec, params, arr = load_case("soft_d40_conddense")
    flow = build_flow(spec, params, device="cpu", precision="fp32")
    flow.log_prob(arr["x"]); flow.backward(arr["x"]); flow._forward(arr["z0"]); flow.sample([3])
    assert sorted(flow._programs, key=str) == [("backward", False), ("backward", True), ("forward", False), ("forward", True)]
    assert rel_err(flow.log_prob(arr["x"], arr["ctx"]), arr["lp32_ctx"]) < 2e-5
    hard = build_flow(dict(spec, soft_training=False), params, device="cpu", precision="fp32")
    assert rel_err(hard.backward(arr["x"]), arr["z32"]) < 5e-5          # without soft training: never a context
    assert rel_err(hard.log_prob(arr["x"]), arr["lp32"]) > 1e-4         # ... so log_prob lacks the context layer's bias


def test_additive_affine_nn(fake_ops):
    """networks.AdditiveAffineNN (networks.py:14-38): [loc, 0] from a DenseNN under `loc_fnc`."""
    import usflows_b200 as U
    torch.manual_seed(1)
    net = U.AdditiveAffineNN(10, [16, 16], 6)
    assert list(net.state_dict()) == [f"loc_fnc.layers.{i}.{n}" for i in range(3) for n in ("weight", "bias")]
    x = torch.randn(5, 10)
    loc, log_scale = net(x)
    L = net.loc_fnc.layers
    want = L[2](torch.relu(L[1](torch.relu(L[0](x))))).detach()
    assert rel_err(loc, want) < 1e-6 and log_scale.shape == loc.shape and not log_scale.any()


def test_bottleneck_conv_conditioner_runs_layer_by_layer(fake_ops):
    """networks.BottleneckConv (networks.py:754-824) as the coupling conditioner of an image-shaped USFlow: the reference's
    outputs through the layer-by-layer route (no fused launch program exists for it), reference state-dict keys."""
    import usflows_b200 as U
    spec, params, arr = load_case("img_bottleneck_c4_5x4")
    flow = build_flow(spec, params, device="cpu", precision="fp32")
    assert flow._layer_route and not flow._programs
    assert rel_err(flow.log_prob(arr["x"]), arr["lp32"]) < 2e-5
    assert rel_err(flow.backward(arr["x"]), arr["z32"]) < 5e-5
    assert rel_err(flow._forward(arr["z0"]), arr["y32"]) < 5e-5
    assert tuple(flow.sample([3]).shape) == (3, 4, 5, 4) and not flow._programs
    assert rel_err(flow.log_prob_host(arr["x"]), arr["lp32"]) < 2e-5
    net = U.BottleneckConv(4, None, None, [4, 5, 4], c_hidden=6)
    assert sorted(net.state_dict()) == sorted(f"{g}.{i}.{n}" for g in ("in_convolutions", "linear_layers", "out_convolutions")
                                               for i in (0, 1) for n in ("weight", "bias"))
    x = torch.rand(3, 4, 5, 4)
    want = x
    for conv in net.in_convolutions:
        want = torch.relu(conv(want))
    want = want.view(3, -1)
    for lin in net.linear_layers:
        want = torch.relu(lin(want))
    want = want.view(3, 1, 5, 4)
    for conv in net.out_convolutions:
        want = torch.relu(conv(want))
    assert net(x).shape == (3, 4, 5, 4) and rel_err(net(x), want.detach()) < 1e-5
    # gradients of the training pass against autograd through the oracle
    from usflows_b200 import training
    from oracle import flow_oracle as O
    loss = -training.log_prob_autograd(flow, arr["x"][:12]).mean()
    loss.backward()
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in params.items()}
    want_loss = -O.flow_log_prob(arr["x"][:12], spec, p).mean()
    want_loss.backward()
    assert abs(float(loss.detach()) - float(want_loss.detach())) <= 2e-5 * max(1.0, abs(float(want_loss.detach())))
    got = dict(flow.named_parameters())
    checked = 0
    for key in got:
        if "conditioner" in key:
            g, w = got[key].grad, p[key].grad
            assert g is not None and w is not None, key
            assert float((g - w).abs().max()) <= 2e-4 * max(1.0, float(w.abs().max())), key
            checked += 1
    assert checked == 24


def test_reference_accessors(fake_ops):
    """Small pieces of the reference's surface around the path: `LUTransform.to_linear` (transforms.py:1365-1369),
    `HouseholderTransform._construct_householder_permutation` (:795-809), the Transform properties (:195-202, 247-251),
    `DistributionModule.distribution` (distributions.py:129-138), `RadialDistribution.log_delta_volume` (:514-549), a
    radius distribution's own `log_prob` / `sample`."""
    import math
    import usflows_b200 as U
    from oracle import flow_oracle as O
    torch.manual_seed(3)
    lu = U.LUTransform(6, prior_scale=1.0)
    lin = lu.to_linear()
    W = (lu.L_raw.tril(-1) + torch.eye(6)) @ lu.U_raw.triu()
    assert rel_err(lin.forth.weight.detach(), torch.linalg.inv(W.double())) < 1e-5
    assert rel_err(lin.back.weight.detach(), W.detach()) < 1e-6
    assert abs(float(lin.log_abs_det_jacobian(None, None)) + float(lu.log_abs_det_jacobian(None, None))) < 1e-5
    hh = U.HouseholderTransform(5, nvs=2)
    want = hh.w_0.detach().double()
    for v in hh.vk_householder.detach().double():
        want = want @ (torch.eye(5, dtype=torch.float64) - 2 * torch.outer(v, v) / v.dot(v))
    assert rel_err(hh._construct_householder_permutation(), want) < 1e-6
    perm = U.Permute(torch.randperm(6))
    assert perm.with_cache(1) is perm and perm.domain.event_dim == 1 and perm.codomain.event_dim == 1
    lap = U.Laplace(torch.zeros(4), 2 * torch.ones(4))
    d = lap.distribution
    assert isinstance(d, torch.distributions.Independent) and tuple(d.event_shape) == (4,)
    x = torch.randn(7, 4)
    assert rel_err(lap.log_prob(x), d.log_prob(x).detach()) < 1e-6
    assert abs(float(U.Normal(torch.zeros(3), torch.ones(3)).distribution.entropy()) - 3 * 0.5 * math.log(2 * math.pi * math.e)) < 1e-5
    rd = U.RadialDistribution(torch.zeros(9), U.LogNormal(torch.ones(1), torch.ones(1)), p=2.0)
    r = torch.tensor([0.5, 1.0, 3.0], dtype=torch.float64)
    for p in (1.0, 2.0, math.inf):
        assert rel_err(rd.log_delta_volume(p, r), O.radial_log_delta_volume(p, r, 9)) < 1e-12
    chi = U.Chi(5.0, 1.5)                      # cdf / entropy as the reference's Chi states them (distributions.py:98-115)
    v = torch.tensor([0.5, 2.0, 4.0])
    ref_chi2 = torch.distributions.Chi2(torch.tensor(5.0))
    assert torch.allclose(chi.cdf(v), ref_chi2.cdf((v / 1.5) ** 2))
    assert abs(float(chi.entropy()) - float(ref_chi2.entropy() / 2 + math.log(2) + math.log(1.5))) < 1e-6
    # a radius distribution on its own: log f_R(r) and draws of R
    for nd, ref in ((U.Chi(5.0, 1.5), O._Chi(torch.tensor(5.0), 1.5)),
                    (U.GammaMM(torch.tensor([2.0, 3.0]), torch.tensor([1.0, 0.5]), torch.tensor([0.2, -0.1])), None),
                    (U.LogNormal(torch.tensor([0.3]), torch.tensor([0.4])), torch.distributions.LogNormal(0.3, 0.4))):
        rr = torch.rand(11, 1) * 3 + 0.1
        got = nd.log_prob(rr)
        assert got.shape == (11,)
        if ref is None:
            sp = torch.nn.functional.softplus
            ref = torch.distributions.MixtureSameFamily(torch.distributions.Categorical(logits=nd.mixture_logits.detach()),
                                                        torch.distributions.Gamma(sp(nd.concentration_unconstrained.detach()),
                                                                                  sp(nd.rate_unconstrained.detach())))
            want = ref.log_prob(rr[:, 0])
        else:
            want = ref.log_prob(rr)[:, 0]
        assert rel_err(got, want) < 1e-5
        s = nd.sample([50])
        assert s.shape == (50, 1) and bool((s > 0).all())
