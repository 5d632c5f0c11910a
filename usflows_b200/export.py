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
        for net in getattr(flow, "_cond_dense_nets", ()):      # zero context in log_prob / sample of a soft USFlow only
            net.zero_context_default = bool(getattr(flow, "_zero_ctx", False)) and mode in ("log_prob", "sample")
        if mode not in ("log_prob", "backward", "forward", "sample"):
            raise ValueError(f"Unknown export mode {mode}")
        self.mode = mode
        self.event_shape = tuple(int(v) for v in flow._event_shape())
        self.plans = {}                      # layer index -> op list of a ConvNet / ConvNet2D conditioner
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
            if isinstance(layer, (T.BlockAffineTransform, T.Bijective1x1Conv2d)):
                W, Winv, b, ladj = _affine_tensors(getattr(layer, "block_transform", layer))
                self.kinds.append("affine")
                for t in (W, Winv, b, ladj * layer.n_blocks):
                    reg(t)
            elif isinstance(layer, T.AffineTransform):
                W, Winv, b, ladj = _affine_tensors(layer)
                self.kinds.append("affine")
                for t in (W, Winv, b, ladj * getattr(layer, "n_blocks", 1)):     # BlockLUTransform: per block
                    reg(t)
            elif type(layer) is T.MaskedCoupling and not hasattr(layer.conditioner, "layers"):
                # the reference's own conditioners: networks.ConvNet (vector branch) / networks.ConvNet2D
                self.kinds.append("coupling_net")
                reg(layer.mask.reshape(1, *layer.mask.shape[-len(self.event_shape):]))
                self.plans[len(self.kinds) - 1] = self._net_plan(layer.conditioner, reg, start)
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
                from .nn import mlp_layers
                for lin in mlp_layers(layer.conditioner):            # ConditionalDenseNN: without / with zero context
                    reg(lin.weight)
                    reg(lin.bias)
            elif isinstance(layer, T.ScaleTransform):
                self.kinds.append("scale")
                reg(layer.scale.reshape(self.event_shape))
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
        sp = torch.nn.functional.softplus
        self.register_buffer("loc", base.loc.detach().clone().to(torch.float32).reshape(self.event_shape))
        if type(base).__name__ == "RadialDistribution":            # distributions.py:501-549
            nd = base.norm_distribution
            self.base_kind, self.p, self.dv_const = "radial", base.p, base.log_delta_volume_const()
            if hasattr(nd, "_lognormals"):                         # log-normal family (single or mixture)
                self.norm_kind = "lognormal"
                logits, mu, sg = (t.detach().clone() for t in nd._lognormals())
                self.register_buffer("r_logw", torch.log_softmax(logits, 0))
                self.register_buffer("r_mu", mu)
                self.register_buffer("r_sigma", sg)
            else:                                                  # (generalised) Gamma family
                self.norm_kind = "gammamm"
                logits, conc, rate, scale, power = (None if t is None else t.detach().clone() for t in nd._mixture())
                self.register_buffer("r_logw", torch.log_softmax(logits, 0))
                self.register_buffer("r_conc", conc)
                self.register_buffer("r_rate", rate)
                self.register_buffer("r_scale", scale)         # not None: R = scale S^(1 / power)  (Chi, Weibull, HalfNormal)
                self.register_buffer("r_power", power)
        else:
            self.base_kind = "laplace" if type(base).__name__ == "Laplace" else "normal"
            raw = base.scale_unconstrained.detach()
            scale = sp(raw.expand_as(base.loc) if raw.dim() == 0 else raw).reshape(self.event_shape)
            self.register_buffer("scale", scale.clone().to(torch.float32))

    @staticmethod
    def _net_plan(net, reg, start):
        """Op list of a ConvNet (vector) / ConvNet2D conditioner; tensors are registered in order, ops refer to them by
        position (networks.py:222-245, 287-307 / 61-121, 441-494)."""
        if not hasattr(net, "_describe"):
            raise NotImplementedError(f"usflows_b200: the exportable reference semantics do not cover {type(net).__name__} "
                                      "conditioners")
        d = net._describe()
        conv = "conv1" in (d["blocks"][0] if d["blocks"] else {}) or hasattr(net, "kernel_size")
        plan, n = [], [1]                            # tensor 0 of the layer is the mask

        def lin(m, tag):
            reg(m.weight)
            reg(m.bias)
            plan.append((tag, n[0], int(m.dilation[0]) if hasattr(m, "dilation") else 1))
            n[0] += 2

        lin(d["first"], "conv" if conv else "lin")
        for blk in d["blocks"]:
            if conv:
                if blk["gated"]:
                    plan.append(("save", 0, 0))
                    plan.append(("relu", 0, 0))
                    lin(blk["conv1"], "conv")
                    plan.append(("relu", 0, 0))
                    lin(blk["conv2"], "conv")
                    if blk.get("proj") is not None:          # GatedConvND whose width changes (networks.py:186-201)
                        reg(blk["proj"].weight)
                        reg(blk["proj"].bias)
                        plan.append(("proj_conv", n[0], 0))
                        n[0] += 2
                    plan.append(("gate", 0, 0))
                else:
                    lin(blk["conv1"], "conv")
                plan.append(("relu", 0, 0))
                if blk["ln"] is not None:
                    reg(blk["ln"].gamma)
                    reg(blk["ln"].beta)
                    plan.append(("ln_channels", n[0], float(blk["ln"].eps)))
                    n[0] += 2
            else:
                if blk["gated"]:
                    plan.append(("save", 0, 0))
                    plan.append(("relu", 0, 0))
                    lin(blk["lin1"], "lin")
                    plan.append(("relu", 0, 0))
                    lin(blk["lin2"], "lin")
                    if blk["proj"] is not None:
                        reg(blk["proj"].weight)
                        reg(blk["proj"].bias)
                        plan.append(("proj", n[0], 0))
                        n[0] += 2
                    plan.append(("gate", 0, 0))
                else:
                    plan.append(("relu", 0, 0))
                    lin(blk["lin1"], "lin")
                if blk["ln"] is not None:
                    reg(blk["ln"].weight)
                    reg(blk["ln"].bias)
                    plan.append(("ln", n[0], float(blk["ln"].eps)))
                    n[0] += 2
        lin(d["last"], "conv" if conv else "lin")
        return plan

    @staticmethod
    def _run_plan(plan, ts, h):
        F = torch.nn.functional
        saved = None
        for op, i, arg in plan:
            if op == "lin":
                h = F.linear(h, ts[i], ts[i + 1])
            elif op == "conv":
                h = F.conv2d(h, ts[i], ts[i + 1], padding="same", dilation=arg)
            elif op == "relu":
                h = torch.relu(h)
            elif op == "save":
                saved = h
            elif op == "proj":
                saved = F.linear(saved, ts[i], ts[i + 1])
            elif op == "proj_conv":
                saved = F.conv2d(saved, ts[i], ts[i + 1])
            elif op == "gate":
                val, gate = h.chunk(2, dim=1)
                h = saved + val * torch.sigmoid(gate)
            elif op == "ln":
                h = F.layer_norm(h, (h.shape[1],), ts[i], ts[i + 1], arg)
            else:                                    # LayerNormChannels, networks.py:53-58
                mean = h.mean(dim=1, keepdim=True)
                var = h.var(dim=1, unbiased=False, keepdim=True)
                h = (h - mean) / torch.sqrt(var + arg) * ts[i] + ts[i + 1]
        return h

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
            if len(self.event_shape) == 3:                          # 1x1 convolution over the channels (transforms.py:904-962)
                conv = torch.nn.functional.conv2d
                y = conv(x, W[:, :, None, None], b) if to_data else conv(x - b.view(1, -1, 1, 1), Winv[:, :, None, None])
            else:
                y = torch.nn.functional.linear(x, W, b) if to_data else torch.nn.functional.linear(x - b, Winv)
            return y, sign * ladj                                   # transforms.py:913-980
        if kind == "coupling_net":
            m = ts[0]
            t = (1 - m) * self._run_plan(self.plans[i], ts, x * m)
            return (x + t if to_data else x - t), None              # transforms.py:277-306, 316-326
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
            if self.base_kind == "radial":
                raise NotImplementedError("usflows_b200: the exported sampler covers Laplace / Normal bases")
            if self.base_kind == "laplace":
                u = torch.rand_like(x) - 0.5
                e = -torch.sign(u) * torch.log1p(-2.0 * u.abs())
            else:
                e = torch.randn_like(x)
            return self._to_data(self.loc + self.scale * e)
        z, total = self._to_latent(x)      # flows.py:234-245
        ev = tuple(range(1, z.dim()))
        if self.base_kind == "radial":     # distributions.py:501-549
            v = z - self.loc
            r = v.abs().sum(ev) if self.p == 1.0 else v.pow(2).sum(ev).sqrt() if self.p == 2.0 else v.abs().amax(ev)
            logr = r.log()
            if self.norm_kind == "lognormal":
                t = self.r_logw - ((logr[:, None] - self.r_mu) ** 2) / (2 * self.r_sigma ** 2) - self.r_sigma.log() \
                    - 0.5 * math.log(2 * math.pi)
                lpr = torch.logsumexp(t, -1) - logr
            else:
                t = self.r_logw + self.r_conc * self.r_rate.log() - torch.lgamma(self.r_conc)
                if self.r_scale is None:
                    t = t + (self.r_conc - 1) * logr[:, None] - self.r_rate * r[:, None]
                else:
                    lu = logr[:, None] - self.r_scale.log()
                    t = t + self.r_power.log() - self.r_scale.log() + (self.r_conc * self.r_power - 1) * lu \
                        - self.r_rate * torch.exp(self.r_power * lu)
                lpr = torch.logsumexp(t, -1)
            return lpr - (self.dv_const + (self.loc.numel() - 1) * logr) - total
        if self.base_kind == "laplace":
            lp = -torch.log(2 * self.scale) - (z - self.loc).abs() / self.scale
        else:
            lp = -((z - self.loc) ** 2) / (2 * self.scale ** 2) - self.scale.log() - 0.5 * math.log(2 * math.pi)
        return lp.sum(ev) - total


def to_onnx(flow, path: str, export_mode: str = "log_prob", **export_kwargs) -> None:
    """`Flow.to_onnx` (flows.py:212-223): the frozen reference-semantics module, traced on one base-shaped row.
    Needs the `onnx` package like any `torch.onnx.export`."""
    module = ReferenceSemantics(flow, export_mode).cpu().eval()
    dummy = torch.zeros(1, *module.event_shape, dtype=torch.float32)
    export_kwargs.setdefault("input_names", ["x"])
    export_kwargs.setdefault("output_names", [export_mode])
    export_kwargs.setdefault("dynamic_axes", {"x": {0: "rows"}, export_mode: {0: "rows"}})
    export_kwargs.setdefault("dynamo", False)
    torch.onnx.export(module, (dummy,), path, **export_kwargs)
