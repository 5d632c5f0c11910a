"""CPU ORACLE for the USFlows batched flow-evaluation hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain restatement of the reference's algorithm (aai-institute/USFlows,
`src/usflows/{flows,transforms,distributions,utils}.py`) for `log_prob`, `backward`, `_forward`/`sample`
through a flat (1-D `in_dims`) or image-shaped USFlow stack.  Every function cites the reference lines it
follows.  It is written with torch *CPU* tensor ops because the reference itself is eager PyTorch: the same
ATen calls in the same order (including the reference's per-call O(d^3) weight re-preparation), so that
fp32 results agree with the reference to rounding and its run time is a faithful CPU baseline.

Who may import this: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg, and there only as the checker / the CPU baseline -- never the product package `usflows_b200`.

Parity pin: validated against the real reference (imported through `oracle/ref_shim`, a test-only stand-in
for the missing `pyro-ppl`) by `oracle/make_golden.py`, which also writes the committed fixtures in
`tests/golden/`; `tests/test_oracle.py` re-checks the oracle against those fixtures and against the
reference's own four known-answer tests (`tests/veriflow/transforms_test.py:5-67`).
Third-party arithmetic not under /root/reference: `pyro.nn.DenseNN` (pyro-ppl 1.8.6, poetry.lock:3198) --
restated in `dense_nn` below from its published definition (Linear/ReLU stack, no output activation);
the reference has no test for it, so that single function is "parity unpinned" against real pyro.

The model is described by
  spec  : dict(in_dims, coupling_blocks, hidden_dims, affine_conjugation, lu_transform, householder,
               base ("laplace"|"normal"|"radial" with p (1|2|"inf"), norm ("lognormal"|"gammamm"|"gamma"|"chi"|"chi2"|
                     "halfnormal"|"weibull"|"exponential"|"torchlognormal"|"weibullmm"|"lognormalmm" with n_comp / df /
                     chi_scale / w_scale / w_conc / rate / ln_loc / ln_scale)),
               masktype ("checkerboard"|"channel"),
               conditioner ("densenn" (default) | "convnet" with c_hidden, gating, normalize_layers
                            | "convnet2d" with c_hidden (int), num_layers, kernel_size, gating, normalize_layers: image-shaped
                              in_dims = [C, H, W]
                            | "condconvnet" / "condconvnet2d" / "conddense" (context-conditioned; soft_training, `_context`)
                            | "bottleneck" (networks.BottleneckConv)))
  params: dict name -> torch CPU tensor, keyed exactly like the reference `USFlow.state_dict()`.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------
# layer list construction (reference flows.py:389-491)
# --------------------------------------------------------------------------------------------------
def checkerboard_mask(in_dims: List[int], dtype=torch.float32) -> Tensor:
    """flows.py:494-514: mask = (sum of index coordinates) mod 2, shape (1, *in_dims)."""
    axes = [torch.arange(d, dtype=torch.int32) for d in in_dims]
    idx = torch.stack(torch.meshgrid(*axes, indexing="ij"))
    return torch.fmod(idx.sum(dim=0), 2).to(dtype).view(1, *in_dims)


def channel_mask(in_dims: List[int], dtype=torch.float32) -> Tensor:
    """flows.py:516-536: mask = (first index coordinate) mod 2."""
    axes = [torch.arange(d, dtype=torch.int32) for d in in_dims]
    idx = torch.stack(torch.meshgrid(*axes, indexing="ij"))
    return torch.fmod(idx[0], 2).to(dtype).view(1, *in_dims)


def build_layers(spec: dict) -> List[dict]:
    """Layer descriptors in `Flow.layers` order (flows.py:434-482).

    kinds: "affine" (BlockAffineTransform over Sequential[LU..., Householder?] or a bare LU),
           "coupling" (MaskedCoupling + DenseNN), "inv_affine" (InverseTransform of block i's affine),
           "scale" (ScaleTransform).  `prefix` is the state-dict prefix of the layer's parameters.
    """
    in_dims = list(spec["in_dims"])
    B = spec["coupling_blocks"]
    conj = spec.get("affine_conjugation", False)
    n_lu = spec.get("lu_transform", 1)
    hh = spec.get("householder", 1)
    gen = checkerboard_mask if spec.get("masktype", "checkerboard") == "checkerboard" else channel_mask
    mask = gen(in_dims)
    layers: List[dict] = []
    for _ in range(B):
        n_aff = n_lu + (1 if hh > 0 else 0)
        aff = None
        if n_aff > 0:
            aff = dict(kind="affine", prefix=f"trainable_layers.{len(layers)}.block_transform.",
                       seq=True, n_lu=n_lu, householder=hh)
            layers.append(aff)
        layers.append(dict(kind="coupling", prefix=f"trainable_layers.{len(layers)}.conditioner.", mask=mask))
        if conj and aff is not None:
            layers.append(dict(kind="inv_affine", inner=aff,
                               prefix=f"trainable_layers.{len(layers)}.transform.block_transform."))
        mask = 1 - mask
    layers.append(dict(kind="affine", prefix=f"trainable_layers.{len(layers)}.block_transform.",
                       seq=False, n_lu=1, householder=0))
    layers.append(dict(kind="scale", prefix=f"trainable_layers.{len(layers)}."))
    return layers


# --------------------------------------------------------------------------------------------------
# LUTransform (transforms.py:1178-1379)
# --------------------------------------------------------------------------------------------------
def lu_L(L_raw: Tensor) -> Tensor:
    """transforms.py:1271-1274: L = tril(L_raw, -1) + I."""
    return L_raw.tril(-1) + torch.eye(L_raw.shape[0], dtype=L_raw.dtype)


def lu_U(U_raw: Tensor) -> Tensor:
    """transforms.py:1276-1279: U = triu(U_raw)."""
    return U_raw.triu()


def lu_matrix(L_raw: Tensor, U_raw: Tensor) -> Tensor:
    """transforms.py:1281-1283: L @ U."""
    return torch.matmul(lu_L(L_raw), lu_U(U_raw))


def lu_inverse_matrix(L_raw: Tensor, U_raw: Tensor) -> Tensor:
    """transforms.py:1289-1293: inverse(U) @ inverse(L)  (two separate `torch.inverse` calls)."""
    return torch.matmul(torch.inverse(lu_U(U_raw)), torch.inverse(lu_L(L_raw)))


def lu_ladj(U_raw: Tensor) -> Tensor:
    """transforms.py:1303-1320: sum(log|dU|) with dU = U - triu(U,1) + (1 - I)  (diag()-free form)."""
    U = lu_U(U_raw)
    d = U.shape[0]
    dU = U - U.triu(1) + (torch.ones_like(U) - torch.eye(d, dtype=U.dtype))
    return dU.abs().log().sum()


def lu_forward(x: Tensor, L_raw: Tensor, U_raw: Tensor, bias: Tensor) -> Tensor:
    """transforms.py:1242-1256: F.linear(x, L@U, bias)."""
    return F.linear(x, lu_matrix(L_raw, U_raw), bias)


def lu_backward(y: Tensor, L_raw: Tensor, U_raw: Tensor, bias: Tensor) -> Tensor:
    """transforms.py:1258-1269: ((y - b) @ inv(L)^T) @ inv(U)^T (sequential, not the product)."""
    x = y - bias
    x = F.linear(x, torch.inverse(lu_L(L_raw)))
    return F.linear(x, torch.inverse(lu_U(U_raw)))


def lu_is_feasible(U_raw: Tensor) -> bool:
    """transforms.py:1347-1349."""
    return bool((U_raw.diag() != 0).all())


# --------------------------------------------------------------------------------------------------
# HouseholderTransform (transforms.py:752-872)
# --------------------------------------------------------------------------------------------------
def householder_matrix(vk: Tensor, w_0: Tensor) -> Tensor:
    """transforms.py:795-809: w = w_0 @ prod_k (I - 2 v_k v_k^T / v_k.v_k)."""
    d = w_0.shape[0]
    w = w_0
    for v in vk:
        w = torch.mm(w, torch.eye(d, dtype=w.dtype) - 2 * torch.ger(v, v) / torch.dot(v, v))
    return w


# --------------------------------------------------------------------------------------------------
# SequentialAffineTransform / bare LU behind BlockAffineTransform (transforms.py:1381-1486, 874-1029)
# --------------------------------------------------------------------------------------------------
def _affine_parts(layer: dict, params: Dict[str, Tensor]):
    """Per-sub-transform (matrix, inverse_matrix, bias, ladj) in `transforms` order."""
    p = layer["prefix"]
    parts = []
    if not layer["seq"]:
        L, U, b = params[p + "L_raw"], params[p + "U_raw"], params[p + "bias_vector"]
        return [(lambda: lu_matrix(L, U), lambda: lu_inverse_matrix(L, U), b, lambda: lu_ladj(U))]
    for k in range(layer["n_lu"]):
        q = f"{p}transforms.{k}."
        L, U, b = params[q + "L_raw"], params[q + "U_raw"], params[q + "bias_vector"]
        parts.append((lambda L=L, U=U: lu_matrix(L, U), lambda L=L, U=U: lu_inverse_matrix(L, U), b,
                      lambda U=U: lu_ladj(U)))
    if layer["householder"] > 0:
        q = f"{p}transforms.{layer['n_lu']}."
        vk, w0 = params[q + "vk_householder"], params[q + "w_0"]
        zero = torch.zeros(w0.shape[0], dtype=w0.dtype)
        parts.append((lambda: householder_matrix(vk, w0),
                      lambda: householder_matrix(vk, w0).transpose(0, 1).contiguous(),  # :864-868
                      zero, lambda: torch.zeros((), dtype=w0.dtype)))                    # ladj = 0 (:760)
    return parts


def affine_matrix(layer: dict, params) -> Tensor:
    """transforms.py:1457-1462: M = I; M = M @ A_k.matrix() in order (bare LU: :1281)."""
    parts = _affine_parts(layer, params)
    if not layer["seq"]:
        return parts[0][0]()
    d = parts[0][2].shape[0]
    M = torch.eye(d, dtype=parts[0][2].dtype)
    for mat, _, _, _ in parts:
        M = torch.matmul(M, mat())
    return M


def affine_inverse_matrix(layer: dict, params) -> Tensor:
    """transforms.py:1464-1469: M = I; M = M @ A_k.inverse_matrix() in reverse order."""
    parts = _affine_parts(layer, params)
    if not layer["seq"]:
        return parts[0][1]()
    d = parts[0][2].shape[0]
    M = torch.eye(d, dtype=parts[0][2].dtype)
    for _, inv, _, _ in parts[::-1]:
        M = torch.matmul(M, inv())
    return M


def affine_bias(layer: dict, params) -> Tensor:
    """transforms.py:1471-1476: b = 0; b = b @ A_k.matrix() + b_k in order (bare LU: bias_vector)."""
    parts = _affine_parts(layer, params)
    if not layer["seq"]:
        return parts[0][2]
    b = torch.zeros_like(parts[0][2])
    for mat, _, bk, _ in parts:
        b = torch.matmul(b, mat()) + bk
    return b


def affine_ladj(layer: dict, params, n_blocks: int = 1) -> Tensor:
    """transforms.py:1429-1446 (sum over sub-transforms) x n_blocks (:964-980)."""
    return sum(p[3]() for p in _affine_parts(layer, params)) * n_blocks


def _view_w(w: Tensor, rank: int) -> Tensor:
    return w.view(w.shape[0], w.shape[1], *([1] * rank))


_CONV = {1: F.linear, 2: F.conv1d, 3: F.conv2d, 4: F.conv3d}


def block_affine_forward(x: Tensor, layer: dict, params, in_dims) -> Tensor:
    """transforms.py:913-934: global_transform(x, W.view(d,d,1..), b); F.linear for 1-D in_dims, 1x1 conv else."""
    rank = len(in_dims) - 1
    return _CONV[len(in_dims)](x, _view_w(affine_matrix(layer, params), rank), affine_bias(layer, params))


def block_affine_backward(y: Tensor, layer: dict, params, in_dims) -> Tensor:
    """transforms.py:936-962: global_transform(y - b.view(d,1..), W^-1)."""
    rank = len(in_dims) - 1
    w = _view_w(affine_inverse_matrix(layer, params), rank)
    b = affine_bias(layer, params).view(in_dims[0], *([1] * rank))
    return _CONV[len(in_dims)](y - b, w)


# --------------------------------------------------------------------------------------------------
# MaskedCoupling + pyro.nn.DenseNN (transforms.py:254-347; pyro-ppl 1.8.6 pyro/nn/dense_nn.py)
# --------------------------------------------------------------------------------------------------
def dense_nn(x: Tensor, prefix: str, params, n_layers: int) -> Tensor:
    """pyro.nn.DenseNN.forward with param_dims=[d]: h = relu(Linear(h)) for all but the last Linear."""
    h = x
    for j in range(n_layers - 1):
        h = F.relu(F.linear(h, params[f"{prefix}layers.{j}.weight"], params[f"{prefix}layers.{j}.bias"]))
    j = n_layers - 1
    return F.linear(h, params[f"{prefix}layers.{j}.weight"], params[f"{prefix}layers.{j}.bias"])


def convnet_vector(x: Tensor, prefix: str, params, spec: dict) -> Tensor:
    """networks.ConvNet.forward for 1-D in_dims (networks.py:379-389) over the module list built at :287-307:
    Linear(d, h0) -> per hidden width [GatedMLP (:222-245) | Sequential(ReLU, Linear)] [-> LayerNormVector (:205-219)]
    -> Linear(h_last, c_out).  State-dict names `nn.{i}. ...` as the reference's nn.Sequential numbers them."""
    c_hidden = list(spec["c_hidden"])
    gating, normalize = spec.get("gating", True), spec.get("normalize_layers", True)
    h = F.linear(x, params[f"{prefix}nn.0.weight"], params[f"{prefix}nn.0.bias"])
    idx = 1
    for i, out_ch in enumerate(c_hidden):
        in_ch = c_hidden[i - 1] if i > 0 else c_hidden[0]
        q = f"{prefix}nn.{idx}."
        if gating:                                               # GatedMLP.forward, networks.py:237-245
            out = F.linear(F.relu(h), params[q + "net1.1.weight"], params[q + "net1.1.bias"])
            out = F.linear(F.relu(out), params[q + "net1.3.weight"], params[q + "net1.3.bias"])
            val, gate = out.chunk(2, dim=1)
            res = val * torch.sigmoid(gate)
            if in_ch != out_ch:
                h = F.linear(h, params[q + "proj.weight"], params[q + "proj.bias"])
            h = h + res
        else:                                                    # nn.Sequential(nonlinearity, nn.Linear), networks.py:300
            h = F.linear(F.relu(h), params[q + "1.weight"], params[q + "1.bias"])
        idx += 1
        if normalize:                                            # LayerNormVector -> nn.LayerNorm(eps=1e-5)
            q = f"{prefix}nn.{idx}.layernorm."
            h = F.layer_norm(h, (out_ch,), params[q + "weight"], params[q + "bias"], 1e-5)
            idx += 1
    return F.linear(h, params[f"{prefix}nn.{idx}.weight"], params[f"{prefix}nn.{idx}.bias"])


def convnet2d(x: Tensor, prefix: str, params, spec: dict) -> Tensor:
    """networks.ConvNet2D.forward (networks.py:496-506) over the module list built at :441-494: Conv k x k -> per layer
    [GatedConv (:103-121) | Conv k x k] -> nonlinearity -> [LayerNormChannels (:53-58)] -> Conv k x k, padding 'same'.
    State-dict names `nn.{i}. ...` as nn.Sequential numbers them (the parameter-free ReLU modules take an index too)."""
    L = spec["num_layers"]
    gating, normalize = spec.get("gating", True), spec.get("normalize_layers", True)
    h = F.conv2d(x, params[f"{prefix}nn.0.weight"], params[f"{prefix}nn.0.bias"], padding="same")
    idx = 1
    for _ in range(L):
        q = f"{prefix}nn.{idx}."
        if gating:
            out = F.conv2d(F.relu(h), params[q + "net.1.weight"], params[q + "net.1.bias"], padding="same")
            out = F.conv2d(F.relu(out), params[q + "net.3.weight"], params[q + "net.3.bias"], padding="same")
            val, gate = out.chunk(2, dim=1)
            h = h + val * torch.sigmoid(gate)
        else:
            h = F.conv2d(h, params[q + "weight"], params[q + "bias"], padding="same")
        h = F.relu(h)
        idx += 2
        if normalize:
            q = f"{prefix}nn.{idx}."
            mean = h.mean(dim=1, keepdim=True)
            var = h.var(dim=1, unbiased=False, keepdim=True)
            h = (h - mean) / torch.sqrt(var + 1e-5)
            h = h * params[q + "gamma"] + params[q + "beta"]
            idx += 1
    return F.conv2d(h, params[f"{prefix}nn.{idx}.weight"], params[f"{prefix}nn.{idx}.bias"], padding="same")


def convnet_spatial(x: Tensor, prefix: str, params, spec: dict) -> Tensor:
    """networks.ConvNet.forward for in_dims=[C, H, W] (networks.py:390-403) over the module list built at :308-377:
    Conv k x k (C, h0) -> per hidden width [GatedConvND (:142-203; residual projected by a 1x1 convolution when the width
    changes) | Conv k x k] -> nonlinearity -> [LayerNormChannelsND (:122-140)] -> Conv k x k (h_last, C); padding k // 2."""
    c_hidden = list(spec["c_hidden"])
    gating, normalize = spec.get("gating", True), spec.get("normalize_layers", True)
    pad = int(spec.get("kernel_size", 3)) // 2
    h = F.conv2d(x, params[f"{prefix}nn.0.weight"], params[f"{prefix}nn.0.bias"], padding=pad)
    idx = 1
    for i, out_ch in enumerate(c_hidden):
        in_ch = c_hidden[i - 1] if i > 0 else c_hidden[0]
        q = f"{prefix}nn.{idx}."
        if gating:                                               # GatedConvND.forward, networks.py:196-203
            out = F.conv2d(F.relu(h), params[q + "net.1.weight"], params[q + "net.1.bias"], padding=pad)
            out = F.conv2d(F.relu(out), params[q + "net.3.weight"], params[q + "net.3.bias"])
            val, gate = out.chunk(2, dim=1)
            res = val * torch.sigmoid(gate)
            if in_ch != out_ch:
                h = F.conv2d(h, params[q + "proj.weight"], params[q + "proj.bias"])
            h = h + res
        else:
            h = F.conv2d(h, params[q + "weight"], params[q + "bias"], padding=pad)
        h = F.relu(h)
        idx += 2
        if normalize:                                            # LayerNormChannelsND.forward, networks.py:134-140
            q = f"{prefix}nn.{idx}."
            mean = h.mean(dim=1, keepdim=True)
            var = h.var(dim=1, unbiased=False, keepdim=True)
            h = (h - mean) / torch.sqrt(var + 1e-5)
            h = h * params[q + "gamma"] + params[q + "beta"]
            idx += 1
    return F.conv2d(h, params[f"{prefix}nn.{idx}.weight"], params[f"{prefix}nn.{idx}.bias"], padding=pad)


COND_KINDS = {"condconvnet2d": "convnet2d", "condconvnet": "convnet"}


def with_context(x: Tensor, context) -> Tensor:
    """CondConvNet.forward / CondConvNet2D.forward (networks.py:560-600, 643-680): the context ([N, 1]; None = 0) is
    expanded to (N, 1, *spatial) and appended to x as one more channel (feature, for vector inputs)."""
    n = x.shape[0]
    if context is None:
        context = torch.zeros(1, dtype=x.dtype)
    c = context.to(x.dtype).reshape(*context.shape, *([1] * (x.dim() - context.dim()))).expand(n, 1, *x.shape[2:])
    return torch.cat([x, c], dim=1)


def cond_dense_nn(x: Tensor, prefix: str, params, n_hidden: int, context) -> Tensor:
    """networks.ConditionalDenseNN.forward (networks.py:733-749); layers = [input, context, hidden ..., output]
    (:716-724): h = layers[0](x) (+ layers[1](context)); h = f(h); h = f(layer(h)) for the hidden ones; layers[-1](h)."""
    lin = lambda j, h: F.linear(h, params[f"{prefix}layers.{j}.weight"], params[f"{prefix}layers.{j}.bias"])  # noqa: E731
    h = lin(0, x)
    if context is not None:
        h = h + lin(1, context.to(x.dtype))
    h = F.relu(h)
    for j in range(2, n_hidden + 1):
        h = F.relu(lin(j, h))
    return lin(n_hidden + 1, h)


def bottleneck_conv(x: Tensor, prefix: str, params) -> Tensor:
    """networks.BottleneckConv.forward (networks.py:802-824): conv / ReLU twice (down to one channel), view [N, H*W],
    Linear / ReLU twice, view [N, 1, H, W], conv / ReLU twice (the output is rectified too)."""
    spatial = x.shape[2:]
    for j in range(2):
        x = F.relu(F.conv2d(x, params[f"{prefix}in_convolutions.{j}.weight"], params[f"{prefix}in_convolutions.{j}.bias"],
                            padding="same"))
    x = x.view(x.shape[0], -1)
    for j in range(2):
        x = F.relu(F.linear(x, params[f"{prefix}linear_layers.{j}.weight"], params[f"{prefix}linear_layers.{j}.bias"]))
    x = x.view(x.shape[0], 1, *spatial)
    for j in range(2):
        x = F.relu(F.conv2d(x, params[f"{prefix}out_convolutions.{j}.weight"], params[f"{prefix}out_convolutions.{j}.bias"],
                            padding="same"))
    return x


def conditioner(x: Tensor, layer: dict, params, n_layers: int, spec=None) -> Tensor:
    kind = spec.get("conditioner") if spec is not None else None
    if kind == "bottleneck":
        return bottleneck_conv(x, layer["prefix"], params)
    if kind == "conddense":     # USFlow substitutes a zero context [N, 1] when soft_training and none is given (flows.py:559-565)
        ctx = spec.get("_context")     # `backward` / `_forward` hand none over (flows.py:45-67): the context layer is skipped
        if ctx is None and spec.get("_zero_context"):
            ctx = torch.zeros(x.shape[0], 1, dtype=x.dtype)
        return cond_dense_nn(x, layer["prefix"], params, len(spec["hidden_dims"]), ctx)
    if kind in COND_KINDS:      # soft training (flows.py:172-193, 559-565): spec["_context"] is the context of this call
        x = with_context(x, spec.get("_context"))
        kind = COND_KINDS[kind]
    if kind == "convnet2d":
        return convnet2d(x, layer["prefix"], params, spec)
    if kind == "convnet" and len(spec["in_dims"]) == 3:
        return convnet_spatial(x, layer["prefix"], params, spec)
    if kind == "convnet":
        return convnet_vector(x, layer["prefix"], params, spec)
    return dense_nn(x, layer["prefix"], params, n_layers)


def coupling_forward(x: Tensor, layer: dict, params, n_layers: int, spec=None) -> Tensor:
    """transforms.py:277-290: x + (1 - mask) * conditioner(x * mask)."""
    m = layer["mask"].to(x.dtype)
    return x + (1 - m) * conditioner(x * m, layer, params, n_layers, spec)


def coupling_backward(y: Tensor, layer: dict, params, n_layers: int, spec=None) -> Tensor:
    """transforms.py:292-306: y - (1 - mask) * conditioner(y * mask)."""
    m = layer["mask"].to(y.dtype)
    return y - (1 - m) * conditioner(y * m, layer, params, n_layers, spec)


# --------------------------------------------------------------------------------------------------
# Affine (scale-and-shift) coupling -- EXTENSION, spec["coupling"] == "affine".  The reference has no such layer
# (its MaskedCoupling is additive, transforms.py:316-326 returns 0.0), so there is NO reference output to pin this
# against: "parity unpinned" for this layer.  The form is the task's "masked affine coupling" with the clamped
# log-scale of pyro.distributions.transforms.AffineCoupling (pyro-ppl 1.8.6): the conditioner emits [s | t] (2d values),
#   forward  y = x * exp((1-m) s) + (1-m) t,   backward x = (y - (1-m) t) * exp(-(1-m) s),   s = clamp(s, lo, hi),
#   forward log|det J| per row = sum_j (1-m_j) s_j.
# Tests check it by round trip and against the autograd Jacobian in fp64.
# --------------------------------------------------------------------------------------------------
AFFINE_CLIP = (-5.0, 3.0)


def affine_coupling_terms(x: Tensor, layer: dict, params, n_layers: int):
    m = layer["mask"].to(x.dtype)
    st = dense_nn(x * m, layer["prefix"], params, n_layers)
    d = m.numel()
    s = (1 - m) * st[..., :d].clamp(*AFFINE_CLIP)
    t = (1 - m) * st[..., d:]
    return s, t


def affine_coupling_forward(x: Tensor, layer: dict, params, n_layers: int) -> Tensor:
    s, t = affine_coupling_terms(x, layer, params, n_layers)
    return x * torch.exp(s) + t


def affine_coupling_backward(y: Tensor, layer: dict, params, n_layers: int) -> Tensor:
    s, t = affine_coupling_terms(y, layer, params, n_layers)      # the masked features are unchanged by the layer
    return (y - t) * torch.exp(-s)


def affine_coupling_ladj(x_or_y: Tensor, layer: dict, params, n_layers: int) -> Tensor:
    return affine_coupling_terms(x_or_y, layer, params, n_layers)[0].sum(-1)


# --------------------------------------------------------------------------------------------------
# elementwise layers (transforms.py:73-171, 174-251, 417-474)
# --------------------------------------------------------------------------------------------------
def scale_forward(x: Tensor, scale: Tensor) -> Tensor:
    """transforms.py:105-114."""
    return x * scale


def scale_backward(x: Tensor, scale: Tensor) -> Tensor:
    """transforms.py:116-125."""
    return x / scale


def scale_ladj(scale: Tensor) -> Tensor:
    """transforms.py:135-144."""
    return scale.abs().log().sum()


def leaky_relu_forward(x: Tensor, alpha: float = 0.01) -> Tensor:
    """transforms.py:434-443."""
    return F.leaky_relu(x, negative_slope=alpha)


def leaky_relu_backward(y: Tensor, alpha: float = 0.01) -> Tensor:
    """transforms.py:445-454."""
    return F.leaky_relu(y, negative_slope=1 / alpha)


def leaky_relu_ladj_reference(x: Tensor, y: Tensor) -> Tensor:
    """transforms.py:464-474 verbatim semantics: log(y/x).sum() over the WHOLE tensor (batch included)."""
    return torch.log(y / x).sum()


def leaky_relu_ladj_per_row(x: Tensor, alpha: float = 0.01) -> Tensor:
    """Documented deviation: per-row log|det J| = log(alpha) * #{x_j < 0}; equals the
    reference on its own unbatched known-answer test (transforms_test.py:53-67)."""
    return math.log(alpha) * (x < 0).to(x.dtype).sum(dim=-1)


def permute_forward(x: Tensor, perm: Tensor) -> Tensor:
    """transforms.py:213-223: index_select on the last dim."""
    return x.index_select(-1, perm)


def permute_backward(y: Tensor, perm: Tensor) -> Tensor:
    """transforms.py:201-210, 225-232: index_select with the inverse permutation."""
    inv = torch.empty_like(perm, dtype=torch.long)
    inv[perm] = torch.arange(perm.size(0), dtype=torch.long)
    return y.index_select(-1, inv)


# --------------------------------------------------------------------------------------------------
# base distributions (distributions.py:117-159, 199-238; utils.py:3-9)
# --------------------------------------------------------------------------------------------------
def inv_softplus(x: Tensor) -> Tensor:
    """utils.py:3-9."""
    return torch.log(torch.exp(x) - 1)


def base_scale(params) -> Tensor:
    """distributions.py:211-215 / 228-238: scale = softplus(scale_unconstrained) (scalar expands to loc)."""
    s = F.softplus(params["base_distribution.scale_unconstrained"])
    loc = params["base_distribution.loc"]
    return s.expand_as(loc) if s.dim() == 0 else s


def base_log_prob(z: Tensor, spec: dict, params) -> Tensor:
    """distributions.py:150-151 -> torch.distributions.{Laplace,Normal}.log_prob summed over the event dims
    (DIndependent, distributions.py:133-137).
    Laplace: -log(2 s) - |z - mu| / s ;  Normal: -(z-mu)^2/(2 s^2) - log s - log sqrt(2 pi)."""
    loc = params["base_distribution.loc"]
    if spec.get("base") == "radial":
        return radial_log_prob(z, spec, params)
    s = base_scale(params)
    if spec.get("base", "laplace") == "laplace":
        lp = -torch.log(2 * s) - torch.abs(z - loc) / s
    else:
        var = s ** 2
        lp = -((z - loc) ** 2) / (2 * var) - s.log() - math.log(math.sqrt(2 * math.pi))
    return lp.sum(dim=tuple(range(z.dim() - loc.dim(), z.dim()))) if loc.dim() >= 1 else lp


def _radial_p(spec) -> float:
    return math.inf if spec["p"] == "inf" else float(spec["p"])


def radial_norm_distribution(spec: dict, params):
    """The torch distribution the reference's DistributionModule builds for the radius (distributions.py:129-138):
    LogNormal(loc, softplus(scale_unconstrained)) made Independent over its batch dim (:181-197), or
    MixtureSameFamily(Categorical(logits), Gamma(softplus(.), softplus(.))) (:674-707)."""
    q = "base_distribution.norm_distribution."
    if spec["norm"] == "lognormal":
        d = torch.distributions.LogNormal(params[q + "loc"], F.softplus(params[q + "scale_unconstrained"]))
        return torch.distributions.Independent(d, len(d.batch_shape))
    if spec["norm"] == "chi":           # distributions.py:55-115: scale * sqrt(Chi2(df)); log_prob as written at :88-97
        return _Chi(torch.tensor(float(spec["df"]), dtype=params["base_distribution.loc"].dtype), spec.get("chi_scale", 1.0))
    if spec["norm"] == "chi2":          # torch objects passed straight in (experiments/mnist/mnist_digits_minimal_radial_chi2.yaml:63)
        return torch.distributions.Chi2(torch.tensor(float(spec["df"]), dtype=params["base_distribution.loc"].dtype))
    if spec["norm"] == "halfnormal":    # experiments/mnist/mnist_digits_minimal_radialdists.yaml:95
        return torch.distributions.HalfNormal(torch.tensor(float(spec["chi_scale"]), dtype=params["base_distribution.loc"].dtype))
    dt = params["base_distribution.loc"].dtype
    if spec["norm"] == "weibull":       # experiments/mnist/mnist_digits_minimal_radial_weilbul.yaml:63
        return torch.distributions.Weibull(torch.tensor(float(spec["w_scale"]), dtype=dt), torch.tensor(float(spec["w_conc"]), dtype=dt))
    if spec["norm"] == "exponential":   # experiments/mnist/mnist_digits_minimal_radial_exponential.yaml:63
        return torch.distributions.Exponential(torch.tensor(float(spec["rate"]), dtype=dt))
    if spec["norm"] == "torchlognormal":
        return torch.distributions.LogNormal(torch.tensor(float(spec["ln_loc"]), dtype=dt), torch.tensor(float(spec["ln_scale"]), dtype=dt))
    if spec["norm"] in ("weibullmm", "lognormalmm"):   # MixtureModel (distributions.py:730-795, 821-848)
        p0, p1 = params[q + "unconstrained_params.0"], params[q + "unconstrained_params.1"]
        comp = torch.distributions.Weibull(F.softplus(p0), F.softplus(p1)) if spec["norm"] == "weibullmm" \
            else torch.distributions.LogNormal(p0, F.softplus(p1))
        return torch.distributions.MixtureSameFamily(torch.distributions.Categorical(logits=params[q + "mixture_logits"]), comp)
    conc = F.softplus(params[q + "concentration_unconstrained"])
    rate = F.softplus(params[q + "rate_unconstrained"])
    if spec["norm"] == "gamma":         # distributions.py:162-179, made Independent over its batch dim (:129-138)
        d = torch.distributions.Gamma(conc, rate)
        return torch.distributions.Independent(d, len(d.batch_shape))
    return torch.distributions.MixtureSameFamily(torch.distributions.Categorical(logits=params[q + "mixture_logits"]),
                                                 torch.distributions.Gamma(conc, rate), validate_args=False)


class _Chi:
    """distributions.py:55-115, `log_prob` term by term: chi2.log_prob((v / scale)^2) + log(2 v / scale) - log(scale)."""

    def __init__(self, df, scale):
        self.chi2, self.scale = torch.distributions.Chi2(df), scale

    def log_prob(self, value):
        value = value / self.scale
        return self.chi2.log_prob(value ** 2) + torch.log(value * 2) - torch.log(torch.tensor(self.scale, dtype=value.dtype))

    def sample(self, sample_shape=torch.Size()):
        return self.scale * torch.sqrt(self.chi2.sample(sample_shape))


def radial_log_delta_volume(p: float, r: Tensor, dim: int) -> Tensor:
    """distributions.py:514-549, term by term in the reference's order.  The reference's `self.dim` is a 0-dim int64 TENSOR
    (`torch.prod(torch.tensor(loc.shape))`, :364), so `x * self.dim` / `self.dim / 2` are fp32 tensor operations under the
    default dtype and `math.log(self.dim)` a Python float -- restated with the same operand types (the constants then
    round where the reference's do; with a Python int the p = 2 value differed by one fp32 ulp)."""
    dim = torch.tensor(int(dim))
    dimf = dim.to(r.dtype)      # int tensor (*, /) Python float -> the default dtype, which is r's dtype when the reference runs
    if p == 1:
        log_denominator = sum([math.log(i) for i in range(1, dim)])
        return math.log(2) * dimf + torch.log(r) * (dim - 1) - log_denominator
    if p == 2:
        log_numerator = math.log(dim) + (dimf / 2) * math.log(math.pi) + (dim - 1) * torch.log(r)
        return log_numerator - math.lgamma((dimf / 2) + 1)
    return math.log(dim) + dimf * math.log(2) + (dim - 1) * torch.log(r)


def radial_log_prob(x: Tensor, spec: dict, params) -> Tensor:
    """RadialDistribution.log_prob (distributions.py:501-512): r = ||x - loc||_p over the event dims,
    norm_distribution.log_prob(r[..., None])[..., 0] - log_delta_volume(p, r)."""
    loc = params["base_distribution.loc"]
    p = _radial_p(spec)
    x = x - loc
    event_dims = tuple(range(x.dim() - loc.dim(), x.dim()))
    r = x.norm(p=p, dim=event_dims)
    log_prob_norm = radial_norm_distribution(spec, params).log_prob(r.unsqueeze(-1))
    if log_prob_norm.dim() > r.dim():
        log_prob_norm = log_prob_norm.squeeze(-1)
    return log_prob_norm - radial_log_delta_volume(p, r, loc.numel())


def base_sample_from_uniform(u: Tensor, spec: dict, params) -> Tensor:
    """torch.distributions.Laplace.rsample given its uniform draw u in (-1, 1):
    loc - scale * sign(u) * log1p(-|u|).  (Normal: loc + scale * eps with eps ~ N(0,1) passed as `u`.)"""
    loc = params["base_distribution.loc"]
    s = base_scale(params)
    if spec.get("base", "laplace") == "laplace":
        return loc - s * u.sign() * torch.log1p(-u.abs())
    return loc + s * u


# --------------------------------------------------------------------------------------------------
# Flow.log_prob / backward / _forward (flows.py:225-245, 57-67, 45-55)
# --------------------------------------------------------------------------------------------------
def _cast(params, dtype):
    return {k: v.to(dtype) for k, v in params.items()}


def _n_cond_layers(spec) -> int:
    return len(spec.get("hidden_dims", [])) + 1


def layer_forward(x, layer, spec, params):
    in_dims = spec["in_dims"]
    k = layer["kind"]
    if k == "affine":
        return block_affine_forward(x, layer, params, in_dims)
    if k == "inv_affine":                      # InverseTransform.forward -> inner.backward (:362-368)
        return block_affine_backward(x, layer["inner"], params, in_dims)
    if k == "coupling":
        if spec.get("coupling") == "affine":
            return affine_coupling_forward(x, layer, params, _n_cond_layers(spec))
        return coupling_forward(x, layer, params, _n_cond_layers(spec), spec)
    if k == "scale":
        return scale_forward(x, params[layer["prefix"] + "scale"])
    raise ValueError(k)


def layer_backward(y, layer, spec, params):
    in_dims = spec["in_dims"]
    k = layer["kind"]
    if k == "affine":
        return block_affine_backward(y, layer, params, in_dims)
    if k == "inv_affine":                      # InverseTransform.backward -> inner.forward (:370-376)
        return block_affine_forward(y, layer["inner"], params, in_dims)
    if k == "coupling":
        if spec.get("coupling") == "affine":
            return affine_coupling_backward(y, layer, params, _n_cond_layers(spec))
        return coupling_backward(y, layer, params, _n_cond_layers(spec), spec)
    if k == "scale":
        return scale_backward(y, params[layer["prefix"] + "scale"])
    raise ValueError(k)


def layer_ladj(layer, spec, params) -> Tensor:
    """Forward-direction log|det J| of one layer (a data-independent scalar for every USFlow layer)."""
    n_blocks = math.prod(spec["in_dims"][1:])
    k = layer["kind"]
    if k == "affine":
        return affine_ladj(layer, params, n_blocks)
    if k == "inv_affine":                      # transforms.py:386-396
        return -affine_ladj(layer["inner"], params, n_blocks)
    if k == "coupling":                        # transforms.py:316-326
        return torch.zeros((), dtype=next(iter(params.values())).dtype)
    if k == "scale":
        return scale_ladj(params[layer["prefix"] + "scale"])
    raise ValueError(k)


def flow_backward(x: Tensor, spec: dict, params, dtype=torch.float32) -> Tensor:
    """flows.py:57-67: data -> latent."""
    params = _cast(params, dtype)
    x = x.to(dtype)
    for layer in reversed(build_layers(spec)):
        x = layer_backward(x, layer, spec, params)
    return x


def flow_forward(z: Tensor, spec: dict, params, dtype=torch.float32) -> Tensor:
    """flows.py:45-55: latent -> data (the body of `sample`, flows.py:258-263)."""
    params = _cast(params, dtype)
    z = z.to(dtype)
    for layer in build_layers(spec):
        z = layer_forward(z, layer, spec, params)
    return z


def flow_log_prob(x: Tensor, spec: dict, params, dtype=torch.float32) -> Tensor:
    """flows.py:225-245: log_det = zeros(N); for layer in reversed: y = backward(x); log_det -= ladj;
    return base.log_prob(y) + log_det."""
    params = _cast(params, dtype)
    x = x.to(dtype)
    if spec.get("soft_training") and spec.get("_context") is None:    # USFlow.log_prob: implicit context 0 (flows.py:559-565)
        spec = dict(spec, _zero_context=True)
    log_det = torch.zeros(x.shape[0], dtype=dtype)
    for layer in reversed(build_layers(spec)):
        y = layer_backward(x, layer, spec, params)
        if layer["kind"] == "coupling" and spec.get("coupling") == "affine":       # data-dependent log-det (extension)
            log_det = log_det - affine_coupling_ladj(x, layer, params, _n_cond_layers(spec))
        else:
            log_det = log_det - layer_ladj(layer, spec, params)
        x = y
    return base_log_prob(x, spec, params) + log_det


def flow_total_ladj(spec: dict, params, dtype=torch.float32) -> Tensor:
    """Sum over layers of the forward log|det J| (the model constant subtracted in log_prob)."""
    params = _cast(params, dtype)
    return sum(layer_ladj(l, spec, params) for l in build_layers(spec))


def flow_sample_from_base_draws(u: Tensor, spec: dict, params, dtype=torch.float32) -> Tensor:
    """flows.py:247-265 with the base draw made explicit (so both implementations can share it)."""
    p = _cast(params, dtype)
    return flow_forward(base_sample_from_uniform(u.to(dtype), spec, p), spec, params, dtype)


def flow_log_prob_amortised(x: Tensor, spec: dict, params, dtype=torch.float32, prepared=None):
    """Same arithmetic as `flow_log_prob`, but the weight-side products (matrix / inverse_matrix / bias /
    ladj) are computed once and reused: the "amortised" CPU baseline of BASELINE.md section 4.
    Returns (log_prob, prepared)."""
    params = _cast(params, dtype)
    if spec.get("soft_training") and spec.get("_context") is None:
        spec = dict(spec, _zero_context=True)
    layers = build_layers(spec)
    in_dims = spec["in_dims"]
    rank = len(in_dims) - 1
    if prepared is None:
        prepared = {}
        for i, layer in enumerate(layers):
            if layer["kind"] == "affine":
                prepared[i] = (_view_w(affine_inverse_matrix(layer, params), rank),
                               affine_bias(layer, params).view(in_dims[0], *([1] * rank)))
            elif layer["kind"] == "inv_affine":
                prepared[i] = (_view_w(affine_matrix(layer["inner"], params), rank),
                               affine_bias(layer["inner"], params))
        prepared["ladj"] = sum(layer_ladj(l, spec, params) for l in layers)
        if spec.get("base") != "radial":
            prepared["base_scale"] = base_scale(params)
    x = x.to(dtype)
    conv = _CONV[len(in_dims)]
    for i in range(len(layers) - 1, -1, -1):
        layer = layers[i]
        k = layer["kind"]
        if k == "affine":
            w, b = prepared[i]
            x = conv(x - b, w)
        elif k == "inv_affine":
            w, b = prepared[i]
            x = conv(x, w, b)
        elif k == "coupling":
            x = coupling_backward(x, layer, params, _n_cond_layers(spec), spec)
        else:
            x = scale_backward(x, params[layer["prefix"] + "scale"])
    return base_log_prob(x, spec, params) - prepared["ladj"], prepared


# --------------------------------------------------------------------------------------------------
# deterministic synthetic models (the reference's init distributions, transforms.py:1215-1240, 99-103)
# --------------------------------------------------------------------------------------------------
def random_params(spec: dict, seed: int = 0, min_abs_scale: float = 0.1) -> Dict[str, Tensor]:
    """Random parameters with the reference's init *distributions* (not its RNG stream):
    L_raw strict-lower kaiming-uniform + unit diag, U_raw strict-upper kaiming-uniform + diag
    +-exp(N(0, prior/d)) (transforms.py:1215-1235), bias U(+-1/sqrt d) (:1237-1240), conditioner
    nn.Linear default init, scale U(+-1/sqrt(prod in_dims)) (:99-103) clamped to |s| >= min_abs_scale
    (SURVEY 8d), Householder v ~ 0.2 N(0,1) and w_0 a random permutation matrix (:783-793),
    base loc 0 / scale 1.  Keys = reference state-dict keys (incl. the InverseTransform aliases)."""
    g = torch.Generator().manual_seed(seed)
    in_dims = list(spec["in_dims"])
    d0, dtot = in_dims[0], math.prod(in_dims)
    hidden = list(spec.get("hidden_dims", []))
    out: Dict[str, Tensor] = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def lu(prefix):
        bound = math.sqrt(2.0) * math.sqrt(3.0 / d0)
        L = uni((d0, d0), bound).tril(-1) + torch.eye(d0)
        U = uni((d0, d0), bound).triu(1)
        sign = torch.bernoulli(0.5 * torch.ones(d0), generator=g) * 2 - 1
        diag = sign * torch.exp(torch.randn(d0, generator=g) * (1.0 / d0))
        out[prefix + "L_raw"] = L
        out[prefix + "U_raw"] = U + torch.diag(diag)
        out[prefix + "bias_vector"] = uni((d0,), 1 / math.sqrt(d0))

    for layer in build_layers(spec):
        p = layer["prefix"]
        if layer["kind"] == "affine":
            if not layer["seq"]:
                lu(p)
            else:
                for k in range(layer["n_lu"]):
                    lu(f"{p}transforms.{k}.")
                if layer["householder"] > 0:
                    q = f"{p}transforms.{layer['n_lu']}."
                    out[q + "vk_householder"] = 0.2 * torch.randn(layer["householder"], d0, generator=g)
                    w = torch.zeros(d0, d0)
                    w[torch.arange(d0), torch.randperm(d0, generator=g)] = 1.0
                    out[q + "w_0"] = w
        elif layer["kind"] == "coupling" and spec.get("conditioner") == "bottleneck":      # networks.py:776-798
            ch, k, npix = int(spec["c_hidden"]), int(spec.get("kernel_size", 3)), math.prod(in_dims[1:])
            for name, o, i in (("in_convolutions.0", ch, d0), ("in_convolutions.1", 1, ch),
                               ("out_convolutions.0", ch, 1), ("out_convolutions.1", d0, ch)):
                bound = 1 / math.sqrt(i * k * k)
                out[f"{p}{name}.weight"] = uni((o, i, k, k), bound)
                out[f"{p}{name}.bias"] = uni((o,), bound).abs()       # keeps the rectified stack alive
            for j in range(2):
                out[f"{p}linear_layers.{j}.weight"] = uni((npix, npix), 1 / math.sqrt(npix))
                out[f"{p}linear_layers.{j}.bias"] = uni((npix,), 1 / math.sqrt(npix)).abs()
        elif layer["kind"] == "coupling" and spec.get("conditioner") in ("convnet2d", "condconvnet2d"):
            def conv(name, n_out, n_in, k):
                bound = 1 / math.sqrt(n_in * k * k)
                out[f"{p}{name}.weight"] = uni((n_out, n_in, k, k), bound)
                out[f"{p}{name}.bias"] = uni((n_out,), bound)

            ch, k = int(spec["c_hidden"]), int(spec.get("kernel_size", 3))
            conv("nn.0", ch, d0 + (spec["conditioner"] == "condconvnet2d"), k)       # + the context channel
            idx = 1
            for _ in range(spec["num_layers"]):
                if spec.get("gating", True):
                    conv(f"nn.{idx}.net.1", ch, ch, k)
                    conv(f"nn.{idx}.net.3", 2 * ch, ch, 1)
                else:
                    conv(f"nn.{idx}", ch, ch, k)
                idx += 2
                if spec.get("normalize_layers", True):
                    out[f"{p}nn.{idx}.gamma"] = 1 + uni((1, ch, 1, 1), 0.2)
                    out[f"{p}nn.{idx}.beta"] = uni((1, ch, 1, 1), 0.2)
                    idx += 1
            conv(f"nn.{idx}", d0, ch, k)
        elif layer["kind"] == "coupling" and spec.get("conditioner") in ("convnet", "condconvnet") and len(in_dims) == 3:
            def conv(name, n_out, n_in, k):
                bound = 1 / math.sqrt(n_in * k * k)
                out[f"{p}{name}.weight"] = uni((n_out, n_in, k, k), bound)
                out[f"{p}{name}.bias"] = uni((n_out,), bound)

            ch, k = [int(c) for c in spec["c_hidden"]], int(spec.get("kernel_size", 3))
            conv("nn.0", ch[0], d0 + (spec["conditioner"] == "condconvnet"), k)
            idx = 1
            for i, oc in enumerate(ch):
                ic = ch[i - 1] if i > 0 else ch[0]
                if spec.get("gating", True):
                    conv(f"nn.{idx}.net.1", oc, ic, k)
                    conv(f"nn.{idx}.net.3", 2 * oc, oc, 1)
                    if ic != oc:
                        conv(f"nn.{idx}.proj", oc, ic, 1)
                else:
                    conv(f"nn.{idx}", oc, ic, k)
                idx += 2
                if spec.get("normalize_layers", True):
                    out[f"{p}nn.{idx}.gamma"] = 1 + uni((1, oc, 1, 1), 0.2)
                    out[f"{p}nn.{idx}.beta"] = uni((1, oc, 1, 1), 0.2)
                    idx += 1
            conv(f"nn.{idx}", d0, ch[-1], k)
        elif layer["kind"] == "coupling" and spec.get("conditioner") in ("convnet", "condconvnet"):
            def lin(name, n_out, n_in):
                bound = 1 / math.sqrt(n_in)
                out[f"{p}{name}.weight"] = uni((n_out, n_in), bound)
                out[f"{p}{name}.bias"] = uni((n_out,), bound)

            ch = list(spec["c_hidden"])
            lin("nn.0", ch[0], dtot + (spec["conditioner"] == "condconvnet"))
            idx = 1
            for i, oc in enumerate(ch):
                ic = ch[i - 1] if i > 0 else ch[0]
                if spec.get("gating", True):
                    lin(f"nn.{idx}.net1.1", oc, ic)
                    lin(f"nn.{idx}.net1.3", 2 * oc, oc)
                    if ic != oc:
                        lin(f"nn.{idx}.proj", oc, ic)
                else:
                    lin(f"nn.{idx}.1", oc, ic)
                idx += 1
                if spec.get("normalize_layers", True):        # non-trivial affine part so gamma / beta are exercised
                    out[f"{p}nn.{idx}.layernorm.weight"] = 1 + uni((oc,), 0.2)
                    out[f"{p}nn.{idx}.layernorm.bias"] = uni((oc,), 0.2)
                    idx += 1
            lin(f"nn.{idx}", dtot, ch[-1])
        elif layer["kind"] == "coupling" and spec.get("conditioner") == "conddense":     # networks.py:716-724
            shapes = [(hidden[0], dtot), (hidden[0], 1)] + [(hidden[i], hidden[i - 1]) for i in range(1, len(hidden))] \
                + [(dtot, hidden[-1])]
            for j, (o, i) in enumerate(shapes):
                bound = 1 / math.sqrt(i) if j != 1 else 0.5
                out[f"{p}layers.{j}.weight"] = uni((o, i), bound)
                out[f"{p}layers.{j}.bias"] = uni((o,), bound)
        elif layer["kind"] == "coupling":
            dims = [dtot] + hidden + [2 * dtot if spec.get("coupling") == "affine" else dtot]
            for j in range(len(dims) - 1):
                bound = 1 / math.sqrt(dims[j])
                out[f"{p}layers.{j}.weight"] = uni((dims[j + 1], dims[j]), bound)
                out[f"{p}layers.{j}.bias"] = uni((dims[j + 1],), bound)
        elif layer["kind"] == "inv_affine":
            src = layer["inner"]["prefix"]
            for k in [k for k in out if k.startswith(src)]:
                out[p + k[len(src):]] = out[k]
        elif layer["kind"] == "scale":
            s = uni(tuple(in_dims), 1 / math.sqrt(dtot))
            s = torch.where(s.abs() < min_abs_scale, torch.where(s < 0, -min_abs_scale, min_abs_scale) *
                            torch.ones_like(s), s)
            out[p + "scale"] = s
    if spec.get("base") == "radial":
        q = "base_distribution.norm_distribution."
        out["base_distribution.loc"] = uni(tuple(in_dims), 0.1)
        if spec["norm"] == "lognormal":                       # experiments/mnist/mnist.yaml:86-92 (loc 6, scale .35 at d=784)
            out[q + "loc"] = torch.full((1,), 0.5 * math.log(dtot))
            out[q + "scale_unconstrained"] = inv_softplus(torch.full((1,), 0.35))
        elif spec["norm"] == "gamma":
            out[q + "concentration_unconstrained"] = inv_softplus(0.5 + torch.rand(1, generator=g) * math.sqrt(dtot))
            out[q + "rate_unconstrained"] = inv_softplus(0.5 + torch.rand(1, generator=g))
        elif spec["norm"] in ("chi", "chi2", "halfnormal", "weibull", "exponential", "torchlognormal"):
            pass                                              # no learnable radius parameters (distributions.py:55-75)
        elif spec["norm"] in ("weibullmm", "lognormalmm"):
            K = int(spec.get("n_comp", 4))
            if spec["norm"] == "weibullmm":                   # scale around sqrt(d), shape 1.5 .. 3.5
                out[q + "unconstrained_params.0"] = inv_softplus((0.5 + torch.rand(K, generator=g)) * math.sqrt(dtot))
                out[q + "unconstrained_params.1"] = inv_softplus(1.5 + 2 * torch.rand(K, generator=g))
            else:
                out[q + "unconstrained_params.0"] = 0.5 * math.log(dtot) + 0.5 * torch.rand(K, generator=g)
                out[q + "unconstrained_params.1"] = inv_softplus(0.2 + 0.3 * torch.rand(K, generator=g))
            out[q + "mixture_logits"] = torch.rand(K, generator=g)
        else:                                                 # experiments/synthetic/gaussian_mixture.yaml:84-91
            K = int(spec.get("n_comp", 20))
            out[q + "concentration_unconstrained"] = inv_softplus(0.2 + torch.rand(K, generator=g) * math.sqrt(dtot))
            out[q + "rate_unconstrained"] = inv_softplus(0.5 + torch.rand(K, generator=g))
            out[q + "mixture_logits"] = torch.rand(K, generator=g)
        return out
    out["base_distribution.loc"] = torch.zeros(tuple(in_dims))
    out["base_distribution.scale_unconstrained"] = inv_softplus(torch.ones(tuple(in_dims)))
    return out
