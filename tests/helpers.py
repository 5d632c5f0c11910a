"""Shared test helpers: golden-fixture loading and construction of product flows from an oracle spec."""
import json
import os

import numpy as np
import torch

from oracle import flow_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_CASES = ["c1_d2_laplace", "d2_refinit", "d6_hh_normal", "d5_noconj", "d32_h64", "d100_h50_hh"]
LARGE_CASES = ["c2_d784", "c4_d3072_b2"]
# networks.ConvNet (vector branch) conditioners, Lp-radial bases (LogNormal / GammaMM radius)
EXT_CASES = ["d64_convnet", "d64_convnet_proj_radial2", "d40_convnet_plain_gmm1", "d64_convnet_noln", "d32_radial_inf",
             "d784_radial1_lognormal"]
# image-shaped events [C, H, W] (1x1-convolution BlockAffine, ConvNet2D conditioners, [C, H, W] masks)
IMG_CASES = ["img_c4_4x4", "img_mnist_16x7x7", "img_c6_5x3_plain_channel", "img_c32_4x4_noln"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    spec = json.loads(bytes(z["spec"]).decode())
    params = {k[len("param:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
    if not params:
        params = O.random_params(spec, int(z["seed"]))
    arrays = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files
              if not k.startswith("param:") and k not in ("spec", "seed", "truth64_from_reference")}
    return spec, params, arrays


def build_flow(spec, params, device="cuda", precision=None):
    """usflows_b200.USFlow with the reference state-dict loaded."""
    import usflows_b200 as U
    d = spec["in_dims"][0]
    ev = tuple(spec["in_dims"])
    if spec.get("base") == "radial":
        if spec["norm"] == "lognormal":
            nd = U.LogNormal(torch.ones(1), torch.ones(1))
        else:
            K = spec.get("n_comp", 20)
            nd = U.GammaMM(torch.ones(K), torch.ones(K), torch.ones(K) / K)
        base = U.RadialDistribution(torch.zeros(*ev), nd, p=float("inf") if spec["p"] == "inf" else float(spec["p"]))
    else:
        base = (U.Laplace if spec.get("base", "laplace") == "laplace" else U.Normal)(torch.zeros(*ev), torch.ones(*ev))
    if spec.get("conditioner") == "convnet2d":
        cond_cls = U.ConvNet2D
        cond_args = dict(c_in=d, c_hidden=spec["c_hidden"], num_layers=spec["num_layers"], padding="same",
                         kernel_size=spec.get("kernel_size", 3), normalize_layers=spec.get("normalize_layers", True),
                         gating=spec.get("gating", True))
    elif spec.get("conditioner") == "convnet":
        cond_cls = U.ConvNet
        cond_args = dict(in_dims=[d], c_hidden=list(spec["c_hidden"]), gating=spec.get("gating", True),
                         normalize_layers=spec.get("normalize_layers", True))
    else:
        cond_cls = U.DenseNN
        cond_args = dict(input_dim=d, hidden_dims=list(spec["hidden_dims"]),
                         param_dims=[d, d] if spec.get("coupling") == "affine" else [d])
    flow = U.USFlow(
        base_distribution=base, in_dims=list(spec["in_dims"]),
        coupling_blocks=spec["coupling_blocks"], conditioner_cls=cond_cls,
        conditioner_args=cond_args,
        coupling=spec.get("coupling", "additive"),
        prior_scale=1.0, lu_transform=spec.get("lu_transform", 1), householder=spec.get("householder", 1),
        affine_conjugation=spec.get("affine_conjugation", False), masktype=spec.get("masktype", "checkerboard"),
        precision=precision)
    res = flow.load_state_dict(params, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return flow.to(device)


def rel_err(a, b):
    """max |a - b| / max(max |b|, 1): norm-wise relative error used for every parity statement."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1.0))
