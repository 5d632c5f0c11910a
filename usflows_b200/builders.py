"""Construction of a `USFlow` from a plain spec dict -- the shape vocabulary `bench.py`, `__graft_entry__.smoke()`, the
tools and the tests share (in_dims, coupling_blocks, hidden_dims | conditioner/c_hidden, base, ...; SURVEY 8d lists the
BASELINE configurations in these terms).  Mirrors what the reference's experiment configs do through
`config["model_cfg"]["type"](**params)` (src/usflows/explib/hyperopt.py:101-102)."""
from __future__ import annotations

import torch


def build_flow(spec, params=None, device="cuda", precision=None):
    """usflows_b200.USFlow of the given spec; `params` (a reference-layout state dict) is loaded when given."""
    import usflows_b200 as U
    d = spec["in_dims"][0]
    ev = tuple(spec["in_dims"])
    if spec.get("base") == "radial":
        if spec["norm"] == "lognormal":
            nd = U.LogNormal(torch.ones(1), torch.ones(1))
        elif spec["norm"] == "gamma":
            nd = U.Gamma(torch.ones(1), torch.ones(1))
        elif spec["norm"] == "chi":
            nd = U.Chi(spec["df"], spec.get("chi_scale", 1.0))
        elif spec["norm"] == "chi2":
            nd = torch.distributions.Chi2(torch.tensor(float(spec["df"])))
        elif spec["norm"] == "weibull":
            nd = torch.distributions.Weibull(torch.tensor(float(spec["w_scale"])), torch.tensor(float(spec["w_conc"])))
        elif spec["norm"] == "exponential":
            nd = torch.distributions.Exponential(torch.tensor(float(spec["rate"])))
        elif spec["norm"] == "torchlognormal":
            nd = torch.distributions.LogNormal(torch.tensor(float(spec["ln_loc"])), torch.tensor(float(spec["ln_scale"])))
        elif spec["norm"] == "weibullmm":
            K = spec.get("n_comp", 4)
            nd = U.WeibullMM(torch.ones(K), torch.ones(K), torch.ones(K) / K)
        elif spec["norm"] == "lognormalmm":
            K = spec.get("n_comp", 4)
            nd = U.LogNormalMM(torch.ones(K), torch.ones(K), torch.ones(K) / K)
        elif spec["norm"] == "halfnormal":
            nd = torch.distributions.HalfNormal(torch.tensor(float(spec["chi_scale"])))
        else:
            K = spec.get("n_comp", 20)
            nd = U.GammaMM(torch.ones(K), torch.ones(K), torch.ones(K) / K)
        base = U.RadialDistribution(torch.zeros(*ev), nd, p=float("inf") if spec["p"] == "inf" else float(spec["p"]))
    else:
        base = (U.Laplace if spec.get("base", "laplace") == "laplace" else U.Normal)(torch.zeros(*ev), torch.ones(*ev))
    if spec.get("conditioner") in ("convnet2d", "condconvnet2d"):
        cond_cls = U.ConvNet2D if spec["conditioner"] == "convnet2d" else U.CondConvNet2D
        cond_args = dict(c_in=d, c_hidden=spec["c_hidden"], num_layers=spec["num_layers"], padding="same",
                         kernel_size=spec.get("kernel_size", 3), normalize_layers=spec.get("normalize_layers", True),
                         gating=spec.get("gating", True))
    elif spec.get("conditioner") in ("convnet", "condconvnet"):
        cond_cls = U.ConvNet if spec["conditioner"] == "convnet" else U.CondConvNet
        cond_args = dict(in_dims=list(spec["in_dims"]), c_hidden=list(spec["c_hidden"]), gating=spec.get("gating", True),
                         normalize_layers=spec.get("normalize_layers", True), kernel_size=spec.get("kernel_size", 3))
        if cond_cls is U.CondConvNet:
            cond_args["c_out"] = d
    elif spec.get("conditioner") == "bottleneck":
        cond_cls = U.BottleneckConv
        cond_args = dict(c_in=d, c_hidden_in=None, c_hidden_out=None, in_dims=list(spec["in_dims"]),
                         c_hidden=spec["c_hidden"], kernel_size=spec.get("kernel_size", 3))
    elif spec.get("conditioner") == "conddense":
        cond_cls = U.ConditionalDenseNN
        cond_args = dict(input_dim=d, context_dim=1, hidden_dims=list(spec["hidden_dims"]), out_dim=d)
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
        soft_training=bool(spec.get("soft_training", False)),
        training_noise_prior=torch.distributions.Uniform(1e-20, 0.01) if spec.get("soft_training") else None,
        precision=precision)
    if params is not None:
        res = flow.load_state_dict(params, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
    return flow.to(device)
