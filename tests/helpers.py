"""Shared test helpers: golden-fixture loading and construction of product flows from an oracle spec."""
import json
import math
import os

import numpy as np
import torch

from oracle import flow_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_CASES = ["c1_d2_laplace", "d2_refinit", "d6_hh_normal", "d5_noconj", "d32_h64", "d100_h50_hh"]
LARGE_CASES = ["c2_d784", "c4_d3072_b2"]
# networks.ConvNet (vector branch) conditioners, Lp-radial bases (LogNormal / GammaMM radius)
EXT_CASES = ["d64_convnet", "d64_convnet_proj_radial2", "d40_convnet_plain_gmm1", "d64_convnet_noln", "d32_radial_inf",
             "d784_radial1_lognormal",
             # the other radius distributions of the reference's configurations: its Chi, torch's Chi2 / HalfNormal
             "d32_radial2_chi", "d24_radial1_chi2", "d16_radialinf_halfnormal",
             # torch's Weibull / Exponential / LogNormal objects, the reference's WeibullMM / LogNormalMM radius mixtures
             "d20_radial1_weibull", "d12_radial1_exponential", "d18_radial2_torchlognormal", "d28_radialinf_weibullmm",
             "d30_radial2_lognormalmm",
             "d36_conddense_nocontext"]   # networks.ConditionalDenseNN outside soft training: its context layer is never used
# image-shaped events [C, H, W] (1x1-convolution BlockAffine, ConvNet2D conditioners, [C, H, W] masks)
IMG_CASES = ["img_c4_4x4", "img_mnist_16x7x7", "img_c6_5x3_plain_channel", "img_c32_4x4_noln",
             # networks.ConvNet's convolutional branch (networks.py:308-377) as the conditioner
             "img_convnet_c4_4x4_proj", "img_convnet_c6_5x3_plain", "img_convnet_16x7x7",
             "img_c4_4x4_radial2_gamma"]      # single-Gamma radius (distributions.py:162-179)
# soft training (flows.py:172-193, 559-565): context-conditioned conditioners; fixtures also hold `ctx`, `lp32_ctx`, `lp64_ctx`
SOFT_CASES = ["soft_img_c4_4x4", "soft_img_mnist_16x7x7", "soft_img_convnet_c4_4x4", "soft_d24_convnet",
              "soft_d40_conddense"]      # networks.ConditionalDenseNN: zero context in log_prob, none in backward / _forward


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    spec = json.loads(bytes(z["spec"]).decode())
    params = {k[len("param:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
    if not params:
        params = O.random_params(spec, int(z["seed"]))
    arrays = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files
              if not k.startswith("param:") and k not in ("spec", "seed", "truth64_from_reference")}
    return spec, params, arrays


from usflows_b200.builders import build_flow  # noqa: E402,F401  (re-exported: the tests import it from here)


def rel_err(a, b):
    """max |a - b| / max(max |b|, 1): norm-wise relative error used for every parity statement."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1.0))


def elementwise_err(a, b):
    """(max, 99.9th percentile) of the ELEMENT-wise relative error |a - b| / max(|b|, 1) -- the yardstick SURVEY 8c
    states (`rel_err` above is norm-wise: one large element of b hides absolute errors on the small ones)."""
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    if a.numel() == 0:
        return 0.0, 0.0
    e = (a - b).abs() / b.abs().clamp(min=1.0)
    k = max(1, int(math.ceil(0.999 * e.numel())))
    return float(e.max()), float(e.kthvalue(k).values)


def record_parity(**row):
    """Append one row of achieved parity figures to gpurun_out/parity_elementwise.jsonl (summarised under profiles/)."""
    out = os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_elementwise.jsonl"), "a") as f:
            f.write(json.dumps(row) + "\n")
    except OSError:
        pass


SIMPLIFY_CASES = ["c1_d2_laplace", "d6_hh_normal", "d5_noconj", "d32_h64", "d100_h50_hh", "d64_convnet", "d32_radial_inf",
                  "img_c4_4x4", "img_mnist_16x7x7", "img_c6_5x3_plain_channel",
                  # soft training over ConditionalDenseNN: the reference's simplified flow is a plain `Flow` without the zero
                  # context of USFlow.log_prob, its log-probs differ from the original flow's by the context layer's bias
                  "soft_d40_conddense"]


def load_simplify_case(name):
    """Outputs of the REFERENCE's `flow.simplify()` for a golden case (oracle/make_golden_simplify.py)."""
    z = np.load(os.path.join(GOLDEN, "simplify.npz"))
    meta = json.loads(bytes(z[name + ":meta"]).decode())
    return meta, {k: torch.from_numpy(z[f"{name}:{k}"]) for k in ("lp", "z", "y")}


def layer_kinds(flow):
    """Class names of a flow's layers with the wrapped / block transform, as make_golden_simplify.py records them."""
    return [type(l).__name__ + ("/" + type(l.transform).__name__ if hasattr(l, "transform") else "")
            + ("/" + type(l.block_transform).__name__ if hasattr(l, "block_transform") else "") for l in flow.layers]


def load_layers_golden():
    """Outputs of the reference's Rotation / CompositeRotation / BlockLUTransform (oracle/make_golden_layers.py)."""
    z = np.load(os.path.join(GOLDEN, "layers.npz"))
    return {k: (torch.from_numpy(np.asarray(z[k])) if z[k].shape else float(z[k])) for k in z.files}


def check_standalone_layers(device, tol=1e-5):
    """Rotation, CompositeRotation and BlockLUTransform against the reference's outputs; shared by the emulated-backend
    test (CPU) and the GPU test."""
    import usflows_b200 as U
    G = load_layers_golden()
    x = G["rot:x"].to(device)
    rot = U.Rotation(5, (1, 3), 0.7).to(device)
    y = rot.forward(x)
    assert rel_err(y, G["rot:y"]) <= tol
    assert torch.equal(rot.as_matrix(), G["rot:matrix"])
    assert rel_err(rot.backward(y), G["rot:x"]) <= tol                 # the inverse (the reference's backward is not)
    assert float(rot.log_abs_det_jacobian(x, y)) == 0.0
    comp = U.CompositeRotation([U.Rotation(5, (0, 1), 0.3), U.Rotation(5, (1, 4), -1.1), U.Rotation(5, (2, 0), 2.0)]).to(device)
    yc = comp.forward(x)
    assert rel_err(yc, G["comp:y"]) <= tol
    assert rel_err(comp.as_matrix(), G["comp:as_matrix"]) <= 1e-6      # the reference's (reversed-order) product, literally
    assert rel_err(comp.backward(yc), G["rot:x"]) <= tol
    assert rel_err(yc, G["rot:x"].double() @ comp.matrix().double().cpu().t()) <= tol
    for tag, in_dims in (("blu_flat", [6]), ("blu_img", [4, 3, 5])):
        t = U.BlockLUTransform(in_dims)
        t.load_state_dict({k.split(":param:")[1]: v for k, v in G.items() if k.startswith(tag + ":param:")})
        t = t.to(device)
        xs = G[f"{tag}:x"].to(device)
        assert sorted(t.state_dict()) == ["L_raw", "U_raw", "bias_vector"]
        assert abs(float(t.log_abs_det_jacobian(xs, xs)) - G[f"{tag}:ladj"]) <= 1e-5 * max(1.0, abs(G[f"{tag}:ladj"]))
        assert abs(float(t.log_prior()) - G[f"{tag}:log_prior"]) <= 1e-4 * max(1.0, abs(G[f"{tag}:log_prior"]))
        base = U.Normal(torch.zeros(*in_dims), torch.ones(*in_dims))
        flow = U.Flow(base, [t], device=device)
        assert rel_err(flow._forward(xs), G[f"{tag}:y"]) <= tol
        z = flow.backward(xs)
        assert rel_err(z, G[f"{tag}:z"]) <= tol
        want = torch.distributions.Normal(0.0, 1.0).log_prob(G[f"{tag}:z"].double()).reshape(xs.shape[0], -1).sum(1) \
            - G[f"{tag}:ladj"]
        assert rel_err(flow.log_prob(xs), want) <= tol
        simple = flow.simplify()
        assert rel_err(simple.log_prob(xs), want) <= tol
