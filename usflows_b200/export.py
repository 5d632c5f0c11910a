"""ONNX-exportable reference semantics of a flow (reference flows.py:30-43, 212-223).

`ReferenceSemantics` is a frozen, pure-PyTorch reading of a `Flow`: the same layer algebra in the reference's
operation order, written with traceable tensor ops only, with every weight-side quantity (L.U products, inverses,
Householder products, masks, softplus scales, log-determinants) folded into buffers at construction time.  It exists
for `Flow.to_onnx` and for inspecting a model on any device; it is NOT an execution path of `log_prob` / `sample` /
`backward` / `_forward` -- those run on the sm_100a kernels only and raise on CPU tensors.
"""
from __future__ import annotations

import math
from typing import List

import torch


def _affine_tensors(t):
    from .training import affine_parts
    with torch.no_grad():
        W, Winv, b, ladj = affine_parts(t)
    return W.detach().clone(), Winv.detach().clone(), b.detach().clone(), ladj.detach().clone().reshape(())


class ReferenceSemantics(torch.nn.Module):
    """mode in {"log_prob", "backward", "forward", "sample"} as `Flow.export` (flows.py:30-43)."""

    def __init__(self, flow, mode: str = "log_prob") -> None:
        super().__init__()
        from . import transforms as T
        from .distributions import Independent
        if mode not in ("log_prob", "backward", "forward", "sample"):
            raise ValueError(f"Unknown export mode {mode}")
        self.mode = mode
        self.kinds: List[str] = []
        self.inverted: List[bool] = []
        self.n_tensors: List[int] = []
        count = 0

        def reg(t):
            nonlocal count
            self.register_buffer(f"t{count}", t.detach().clone().to(torch.float32))
            count += 1

        for layer in flow.layers:
            inv = False
            while isinstance(layer, T.InverseTransform):
                inv = not inv
                layer = layer.transform
            self.inverted.append(inv)
            start = count
            if isinstance(layer, T.BlockAffineTransform):
                if len(layer.in_dims) != 1:
                    raise NotImplementedError("usflows_b200: image-shaped in_dims are not built")
                W, Winv, b, ladj = _affine_tensors(layer.block_transform)
                self.kinds.append("affine")
                for t in (W, Winv, b, ladj * layer.n_blocks):
                    reg(t)
            elif isinstance(layer, T.AffineTransform):
                W, Winv, b, ladj = _affine_tensors(layer)
                self.kinds.append("affine")
                for t in (W, Winv, b, ladj):
                    reg(t)
            elif isinstance(layer, T.MaskedCoupling) and not hasattr(layer.conditioner, "layers"):
                raise NotImplementedError("usflows_b200: export of couplings with a ConvNet conditioner is not built")
            elif isinstance(layer, T.MaskedAffineCoupling):
                self.kinds.append("affine_coupling")
                reg(layer.mask.reshape(-1))
                reg(torch.tensor([layer.log_scale_min_clip, layer.log_scale_max_clip]))
                for lin in layer.conditioner.layers:
                    reg(lin.weight)
                    reg(lin.bias)
            elif isinstance(layer, T.MaskedCoupling):
                self.kinds.append("coupling")
                reg(layer.mask.reshape(-1))
                for lin in layer.conditioner.layers:
                    reg(lin.weight)
                    reg(lin.bias)
            elif isinstance(layer, T.ScaleTransform):
                self.kinds.append("scale")
                reg(layer.scale.reshape(-1))
            elif isinstance(layer, T.LeakyReLUTransform):
                self.kinds.append("leaky")
                reg(torch.tensor(float(layer.alpha)))
            elif isinstance(layer, T.Permute):
                self.kinds.append("permute")
                self.register_buffer(f"t{count}", layer.permutation.detach().clone().long())
                count += 1
                self.register_buffer(f"t{count}", layer.inv_permutation.detach().clone().long())
                count += 1
            else:
                raise NotImplementedError(f"usflows_b200: no reference semantics for layer type {type(layer).__name__}")
            self.n_tensors.append(count - start)
        base = flow.base_distribution
        base = base.base_dist if isinstance(base, Independent) else base
        if not hasattr(base, "scale_unconstrained"):
            raise NotImplementedError(f"usflows_b200: export with a {type(base).__name__} base is not built")
        self.base_kind = "laplace" if type(base).__name__ == "Laplace" else "normal"
        raw = base.scale_unconstrained.detach()
        loc = base.loc.detach().reshape(-1)
        scale = torch.nn.functional.softplus(raw.expand_as(base.loc) if raw.dim() == 0 else raw).reshape(-1)
        self.register_buffer("loc", loc.clone().to(torch.float32))
        self.register_buffer("scale", scale.clone().to(torch.float32))

    # -- per-layer algebra: `to_data` = the layer's forward (latent -> data), else its backward -------------------
    def _tensors(self, i: int):
        start = sum(self.n_tensors[:i])
        return [getattr(self, f"t{start + k}") for k in range(self.n_tensors[i])]

    def _layer(self, i: int, x: torch.Tensor, to_data: bool):
        kind, ts = self.kinds[i], self._tensors(i)
        if self.inverted[i]:
            to_data = not to_data
        sign = -1.0 if self.inverted[i] else 1.0
        if kind == "affine":
            W, Winv, b, ladj = ts
            y = torch.nn.functional.linear(x, W, b) if to_data else torch.nn.functional.linear(x - b, Winv)
            return y, sign * ladj                                   # transforms.py:913-980
        if kind == "coupling":
            m = ts[0]
            h = x * m
            n_lin = (len(ts) - 1) // 2
            for j in range(n_lin):
                h = torch.nn.functional.linear(h, ts[1 + 2 * j], ts[2 + 2 * j])
                if j < n_lin - 1:
                    h = torch.relu(h)
            t = (1 - m) * h
            return (x + t if to_data else x - t), None              # transforms.py:277-306, 316-326
        if kind == "affine_coupling":
            m, clip = ts[0], ts[1]
            h = x * m
            n_lin = (len(ts) - 2) // 2
            for j in range(n_lin):
                h = torch.nn.functional.linear(h, ts[2 + 2 * j], ts[3 + 2 * j])
                if j < n_lin - 1:
                    h = torch.relu(h)
            d = m.numel()
            ls = (1 - m) * torch.minimum(torch.maximum(h[:, :d], clip[0]), clip[1])
            t = (1 - m) * h[:, d:]
            y = x * torch.exp(ls) + t if to_data else (x - t) * torch.exp(-ls)
            return y, sign * ls.sum(-1)
        if kind == "scale":
            s = ts[0]
            return (x * s if to_data else x / s), sign * s.abs().log().sum()   # transforms.py:105-144
        if kind == "leaky":
            a = ts[0]
            slope = a if to_data else 1.0 / a
            return torch.where(x >= 0, x, x * slope), None          # data dependent log-det: see log_prob below
        perm, inv_perm = ts
        return x.index_select(-1, perm if to_data else inv_perm), None

    def _to_latent(self, x: torch.Tensor):
        total = torch.zeros((), dtype=x.dtype, device=x.device)
        for i in reversed(range(len(self.kinds))):
            if self.kinds[i] == "leaky":
                raise NotImplementedError("usflows_b200: export of flows with LeakyReLU layers computes no log-det")
            x, ladj = self._layer(i, x, to_data=False)
            if ladj is not None:
                total = total + ladj
        return x, total

    def _to_data(self, z: torch.Tensor) -> torch.Tensor:
        for i in range(len(self.kinds)):
            z, _ = self._layer(i, z, to_data=True)
        return z

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.mode == "backward":
            z = x
            for i in reversed(range(len(self.kinds))):
                z, _ = self._layer(i, z, to_data=False)
            return z
        if self.mode == "forward":
            return self._to_data(x)
        if self.mode == "sample":          # one base draw per row of `x` (shape donor), then latent -> data
            if self.base_kind == "laplace":
                u = torch.rand_like(x) - 0.5
                e = -torch.sign(u) * torch.log1p(-2.0 * u.abs())
            else:
                e = torch.randn_like(x)
            return self._to_data(self.loc + self.scale * e)
        z, total = self._to_latent(x)      # flows.py:234-245
        if self.base_kind == "laplace":
            lp = -torch.log(2 * self.scale) - (z - self.loc).abs() / self.scale
        else:
            lp = -((z - self.loc) ** 2) / (2 * self.scale ** 2) - self.scale.log() - 0.5 * math.log(2 * math.pi)
        return lp.sum(-1) - total


def to_onnx(flow, path: str, export_mode: str = "log_prob", **export_kwargs) -> None:
    """`Flow.to_onnx` (flows.py:212-223): the frozen reference-semantics module, traced on one base-shaped row.
    Needs the `onnx` package like any `torch.onnx.export`."""
    module = ReferenceSemantics(flow, export_mode).cpu().eval()
    d = module.loc.numel()
    dummy = torch.zeros(1, d, dtype=torch.float32)
    export_kwargs.setdefault("input_names", ["x"])
    export_kwargs.setdefault("output_names", [export_mode])
    export_kwargs.setdefault("dynamic_axes", {"x": {0: "rows"}, export_mode: {0: "rows"}})
    export_kwargs.setdefault("dynamo", False)
    torch.onnx.export(module, (dummy,), path, **export_kwargs)
