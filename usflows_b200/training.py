"""Maximum-likelihood training step of a `Flow` (reference `Flow.fit`, src/usflows/flows.py:113-210).

    loss = -mean_batch log_prob(x) - log_prior()          (flows.py:195-198; the prior term is 0)
    loss.backward(); [clip]; optim.step(); is_feasible()   (flows.py:199-205)

The density pass is rebuilt here as an autograd graph whose batch-side contractions -- every `F.linear` of the
affine layers and of the conditioner MLPs, forward AND backward (dX = dY.W, dW = dY^T.X) -- run on the tcgen05
kernels of libusflows_b200.so (`_LinearFn`, tf32-split engine: gradients have no fixed range, so the fp16-split
engine is not used here).  Weight-side work (L@U, triangular inverses, log-dets: O(d^3) once per step) and the
element-wise glue use torch CUDA ops and carry the autograd bookkeeping.

Data parallelism: every rank evaluates its contiguous slice of each global batch with the loss scaled
by 1/global_batch, then ONE all-reduce(sum) of the flat fp32 gradient buffer (NCCL over NVLink on GPUs, gloo in
the CPU tests) reproduces the single-process mean-loss gradient before the (non-linear, sign-based) SophiaG step.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import engine, ops
from .ops import Act, ENGINE_SIMT, ENGINE_TC_3XTF32, pad4

TRAIN_MODE = "fp32_tf32"
TRAIN_ENGINE_NOTE = ("train_engine.TrainEngine where the flow is in its scope (flat LU / DenseNN-coupling stacks: fp16-split "
                     "tcgen05 contractions forward, dX, split-K dW; own weight-side kernels), else torch autograd over the "
                     "library's tf32-split contractions")


# --------------------------------------------------------------------------------------------------
# batch-side contraction with gradients, on the library's kernels
# --------------------------------------------------------------------------------------------------
def _planes(name: str, t: torch.Tensor) -> Act:
    """tf32 hi/lo operand planes of an fp32 matrix (one ingest pass into reusable workspace buffers)."""
    rows, cols = t.shape
    a = Act(rows, cols)
    a.hi = engine._workspace.planes(t.device, name, rows, cols, "hi")
    a.lo = engine._workspace.planes(t.device, name, rows, cols, "lo")
    ops.ingest(t, a)
    return a


def _gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a [M, K] . w [N, K]^T (+ bias) -> fp32 [M, N], fp32-accurate (tf32 split) on tensor cores."""
    M, K = a.shape
    N = w.shape[0]
    a = a.contiguous()
    w = w.contiguous()
    out = torch.empty(M, pad4(N, 4), dtype=torch.float32, device=a.device)[:, :N]
    if M == 0:
        return out
    eng = ENGINE_SIMT if min(N, K) < engine.TC_MIN_DIM else ENGINE_TC_3XTF32
    if eng == ENGINE_SIMT:
        ops.linear(eng, Act(M, K, f32=a), w, None, N, K, bias=bias, out=Act(M, N, f32=out))
    else:
        wa = _planes("tw", w)
        ops.linear(eng, _planes("ta", a), wa.hi, wa.lo, N, K, bias=bias, out=Act(M, N, f32=out))
    return out


def _transposed(t: torch.Tensor) -> torch.Tensor:
    out = torch.empty(t.shape[1], pad4(t.shape[0], 4), dtype=torch.float32, device=t.device)[:, :t.shape[0]]
    ops.transpose(t.contiguous(), out)
    return out


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b with all three contractions (y, dx, dW) on the tcgen05 engine."""

    @staticmethod
    def forward(ctx, x, w, bias):
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return _gemm(x, w.detach(), None if bias is None else bias.detach().contiguous())

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _gemm(dy, _transposed(w.detach()))                    # [M,K] = dY [M,N] . (W^T [K,N])^T
        if ctx.needs_input_grad[1]:
            dw = _gemm(_transposed(dy), _transposed(x.detach()))       # [N,K] = dY^T [N,M] . (X^T [K,M])^T
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dw, db


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _LinearFn.apply(x, w, bias)


# --------------------------------------------------------------------------------------------------
# differentiable layer algebra (weight side in torch, batch side through `linear`)
# --------------------------------------------------------------------------------------------------
def _eye(d, ref):
    return torch.eye(d, dtype=ref.dtype, device=ref.device)


def affine_parts(t, cache: Optional[dict] = None):
    """(W, W^-1, b, log|det W|) of an AffineTransform as differentiable tensors (transforms.py:1271-1320,
    795-809, 1457-1476).  `cache` (one dict per autograd pass) shares the result between a block and the
    `InverseTransform` that wraps the same layer object (affine conjugation): the d^3 products and triangular solves
    are built once per layer and step, and autograd sums the gradients of both uses."""
    if cache is not None:
        hit = cache.get(id(t))
        if hit is None:
            hit = cache[id(t)] = affine_parts(t)
        return hit
    from . import transforms as T
    if isinstance(t, T.LUTransform):
        d = t.dim
        L = t.L_raw.tril(-1) + _eye(d, t.L_raw)
        U = t.U_raw.triu()
        Linv = torch.linalg.solve_triangular(L, _eye(d, L), upper=False, unitriangular=True)
        Uinv = torch.linalg.solve_triangular(U, _eye(d, U), upper=True)
        return L @ U, Uinv @ Linv, t.bias_vector, t.U_raw.diagonal().abs().log().sum()
    if isinstance(t, T.PlaneBijectiveLinearTransform):            # given tensors (transforms.py:618-695); not meant to be trained
        return t.forth.weight, t.back.weight, t.forth.bias, t.ladj
    if isinstance(t, T.Bijective1x1Conv2d):                       # transforms.py:1031-1176, per pixel
        C = t.in_channels
        b = t.forward_conv.bias if t.forward_conv.bias is not None else t.forward_conv.weight.new_zeros(C)
        return t.forward_conv.weight.reshape(C, C), t.inverse_conv.weight.reshape(C, C), b, t.ladj
    if isinstance(t, T.Rotation):                                 # fixed orthogonal maps (transforms.py:476-616)
        R = t._prepared()["matrix"]
        return R, R.t(), R.new_zeros(t.dim), R.new_zeros(())
    if isinstance(t, T.HouseholderTransform):
        W = t.w_0
        for k in range(t.nvs):
            v = t.vk_householder[k]
            W = W - torch.outer(W @ v, v) * (2.0 / torch.dot(v, v))
        return W, W.t(), torch.zeros(t.dim, dtype=W.dtype, device=W.device), W.new_zeros(())
    if isinstance(t, T.SequentialAffineTransform):
        parts = [affine_parts(s) for s in t.transforms]
        M, Minv, b, ladj = parts[0][0], parts[-1][1], parts[0][2], parts[0][3]
        for W, _, bk, l in parts[1:]:
            M = M @ W
            b = b @ W + bk
            ladj = ladj + l
        for _, Winv, _, _ in parts[-2::-1]:
            Minv = Minv @ Winv
        return M, Minv, b, ladj
    raise NotImplementedError(f"usflows_b200: training of {type(t).__name__} is not built")


def _layer_backward(layer, y: torch.Tensor, inverse: bool = False, cache: Optional[dict] = None, geom=None,
                    context: Optional[torch.Tensor] = None):
    """(density direction value, forward log|det J|) of one layer; `inverse` swaps the direction.  For image-shaped events
    `geom = (C, H, W)` and y holds channels-last rows [N*H*W, C] (the layout of image_engine.py): the 1x1 convolution of a
    BlockAffineTransform is then the same `linear` as the flat case, per-element vectors are indexed per pixel."""
    from . import transforms as T
    if isinstance(layer, T.InverseTransform):
        x, ladj = _layer_backward(layer.transform, y, not inverse, cache, geom, context)
        return x, -ladj
    if isinstance(layer, (T.BlockAffineTransform, T.Bijective1x1Conv2d, T.AffineTransform)):
        W, Winv, b, ladj = affine_parts(getattr(layer, "block_transform", layer), cache)
        ladj = ladj * getattr(layer, "n_blocks", 1)    # BlockLUTransform carries its own; a bare affine layer is one block
        if inverse:                                     # the layer's forward: x W^T + b      (transforms.py:913-934)
            return linear(y, W, b), ladj
        return linear(y - b, Winv), ladj                # (y - b) Winv^T                       (transforms.py:936-962)
    if isinstance(layer, T.ScaleTransform):
        ladj = layer.scale.abs().log().sum()
        if geom is not None:                             # scale over [C, H, W], applied per pixel of every image
            C, H, W = geom
            s = layer.scale.reshape(C, H * W).t()
            y3 = y.reshape(-1, H * W, C)
            return (y3 * s if inverse else y3 / s).reshape(-1, C), ladj
        s = layer.scale.reshape(-1)
        return (y * s if inverse else y / s), ladj      # transforms.py:105-125, 135-144
    if isinstance(layer, T.MaskedAffineCoupling):
        if geom is not None:
            raise NotImplementedError("usflows_b200: affine couplings over image-shaped events are not built")
        m = layer.mask.reshape(-1).to(y.dtype)
        h = y * m
        lin = list(layer.conditioner.layers)
        for j, l in enumerate(lin):
            h = linear(h, l.weight, l.bias)
            if j < len(lin) - 1:
                h = torch.relu(h)
        d = m.numel()
        s = (1 - m) * h[:, :d].clamp(layer.log_scale_min_clip, layer.log_scale_max_clip)
        t = (1 - m) * h[:, d:]
        ladj = s.sum(-1)                                               # per row
        return (y * torch.exp(s) + t if inverse else (y - t) * torch.exp(-s)), ladj
    if isinstance(layer, T.MaskedCoupling) and geom is not None:
        C, H, W = geom
        m = layer.mask.reshape(C, H * W).t().to(y.dtype)                # channels-last mask [H*W, C]
        y3 = y.reshape(-1, H * W, C)
        rows = (y3 * m).reshape(-1, C)
        if getattr(layer.conditioner, "layer_route_only", False):      # networks.BottleneckConv
            t = _bottleneck_rows(layer.conditioner, rows, geom).reshape(-1, H * W, C)
        else:
            t = _convnet2d_rows(layer.conditioner, rows, geom, context).reshape(-1, H * W, C)
        t = (1 - m) * t
        return ((y3 + t) if inverse else (y3 - t)).reshape(-1, C), y.new_zeros(())
    if isinstance(layer, T.MaskedCoupling):
        m = layer.mask.reshape(-1).to(y.dtype)
        t = (1 - m) * _conditioner(layer.conditioner, y * m, context)
        return (y + t if inverse else y - t), y.new_zeros(())      # transforms.py:277-306, 316-326
    raise NotImplementedError(f"usflows_b200: training of {type(layer).__name__} is not built")


def _context_rows(context: torch.Tensor, rows: int, like: torch.Tensor) -> torch.Tensor:
    """The per-sample context [N] / [N, 1] as one extra column for `rows` = N * (pixels per sample) rows: the channel a
    conditional network appends, constant over a sample (networks.py:560-600, 643-680)."""
    c = context.to(device=like.device, dtype=like.dtype).reshape(-1, 1)
    if rows % c.shape[0]:
        raise ValueError("context must hold one value per sample")
    return c.repeat_interleave(rows // c.shape[0], dim=0)


def _conditioner(net, h: torch.Tensor, context: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Conditioner network as an autograd graph: contractions on the library's kernels, element-wise glue in torch."""
    from .nn import ConvNet
    if isinstance(net, ConvNet):                      # networks.py:222-245 (GatedMLP), 205-219 (LayerNormVector), 287-307
        with_ctx = context is not None and net.context_channels > 0
        if with_ctx:                                  # CondConvNet, vector branch: [x | context] (networks.py:560-600)
            h = torch.cat([h, _context_rows(context, h.shape[0], h)], dim=1)
        d = net._describe(with_context=with_ctx)
        x = linear(h, d["first"].weight, d["first"].bias)
        for blk in d["blocks"]:
            a = linear(torch.relu(x), blk["lin1"].weight, blk["lin1"].bias)
            if blk["gated"]:
                o = linear(torch.relu(a), blk["lin2"].weight, blk["lin2"].bias)
                val, gate = o.chunk(2, dim=1)
                if blk["proj"] is not None:
                    x = linear(x, blk["proj"].weight, blk["proj"].bias)
                x = x + val * torch.sigmoid(gate)
            else:
                x = a
            if blk["ln"] is not None:
                ln = blk["ln"]
                x = torch.nn.functional.layer_norm(x, ln.normalized_shape, ln.weight, ln.bias, ln.eps)
        return linear(x, d["last"].weight, d["last"].bias)
    if hasattr(net, "_mlp_layers"):                   # ConditionalDenseNN (networks.py:733-749): h = f(L0 x + L1 c), ...
        lin = net._mlp_layers(False)
        h = linear(h, lin[0].weight, lin[0].bias)
        if context is not None:                       # rank-context_dim term: element-wise glue, not a contraction
            c = context.to(device=h.device, dtype=h.dtype).reshape(-1, net.context_dim)
            if c.shape[0] != h.shape[0]:
                raise ValueError("context must hold one row per sample")
            h = h + c @ net.layers[1].weight.t() + net.layers[1].bias
        elif net.zero_context_default:                # the zero context of a soft-training flow (flows.py:559-565)
            h = h + net.layers[1].bias
        for l in lin[1:]:
            h = linear(torch.relu(h), l.weight, l.bias)
        return h
    lin = list(net.layers)
    for j, l in enumerate(lin):
        h = linear(h, l.weight, l.bias)
        if j < len(lin) - 1:
            h = torch.relu(h)
    return h


_GATHER_INDEX = {}


def _gather_index(H: int, W: int, k: int, dil: int, device) -> torch.Tensor:
    """[H*W, k*k] source pixel of every (pixel, tap) of a k x k 'same' convolution; H*W = the zero pixel (padding)."""
    key = (H, W, k, dil, str(device))
    if key not in _GATHER_INDEX:
        hh, ww = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        cols = []
        for kh in range(k):
            for kw in range(k):
                sh, sw = hh + (kh - k // 2) * dil, ww + (kw - k // 2) * dil
                ok = (sh >= 0) & (sh < H) & (sw >= 0) & (sw < W)
                cols.append(torch.where(ok, sh * W + sw, torch.full_like(sh, H * W)).reshape(-1))
        _GATHER_INDEX[key] = torch.stack(cols, dim=1).to(device)
    return _GATHER_INDEX[key]


def _conv_rows(x: torch.Tensor, conv, geom) -> torch.Tensor:
    """nn.Conv2d (stride 1, padding 'same') over channels-last rows [N*H*W, C_in] as gather (torch indexing, carries the
    autograd bookkeeping) + `linear` (contraction forward / dX / dW on the library's kernels)."""
    C, H, W = geom
    cin, k = conv.weight.shape[1], conv.weight.shape[2]
    w = conv.weight.permute(0, 2, 3, 1).reshape(conv.weight.shape[0], -1)        # the usf_im2col column order
    if k == 1:
        return linear(x, w, conv.bias)
    idx = _gather_index(H, W, k, conv.dilation[0], x.device)
    x3 = x.reshape(-1, H * W, cin)
    xp = torch.cat([x3, x3.new_zeros(x3.shape[0], 1, cin)], dim=1)
    cols = xp[:, idx.reshape(-1), :].reshape(-1, k * k * cin)
    return linear(cols, w, conv.bias)


def _convnet2d_rows(net, x: torch.Tensor, geom, context: Optional[torch.Tensor] = None) -> torch.Tensor:
    """networks.ConvNet2D (networks.py:441-506) over channels-last rows as an autograd graph; a conditional network
    (CondConvNet2D / CondConvNet) with a context gets it as one more input channel, else runs without (context 0)."""
    with_ctx = context is not None and getattr(net, "context_channels", 0) > 0
    if with_ctx:
        x = torch.cat([x, _context_rows(context, x.shape[0], x)], dim=1)
    d = net._describe(with_context=with_ctx)
    h = _conv_rows(x, d["first"], geom)
    for blk in d["blocks"]:
        if blk["gated"]:                                  # GatedConv.forward, networks.py:103-121
            o = _conv_rows(torch.relu(_conv_rows(torch.relu(h), blk["conv1"], geom)), blk["conv2"], geom)
            val, gate = o.chunk(2, dim=1)
            if blk.get("proj") is not None:               # GatedConvND whose width changes (networks.py:186-201)
                h = _conv_rows(h, blk["proj"], geom)
            h = h + val * torch.sigmoid(gate)
        else:
            h = _conv_rows(h, blk["conv1"], geom)
        h = torch.relu(h)
        if blk["ln"] is not None:                         # LayerNormChannels, networks.py:53-58 (per pixel = per row)
            ln = blk["ln"]
            h = torch.nn.functional.layer_norm(h, (h.shape[1],), ln.gamma.reshape(-1), ln.beta.reshape(-1), ln.eps)
    return _conv_rows(h, d["last"], geom)


def _bottleneck_rows(net, x: torch.Tensor, geom) -> torch.Tensor:
    """networks.BottleneckConv.forward (networks.py:802-824) over channels-last rows [N*H*W, C]: convolutions down to one
    channel, the flattened pixels [N, H*W] through two Linear layers, convolutions back up; a ReLU after every layer."""
    C, H, W = geom
    h = x
    for conv in net.in_convolutions:
        h = torch.relu(_conv_rows(h, conv, geom))
    h = h.reshape(-1, H * W)                               # one channel left: a sample's rows are its flattened pixels
    for lin in net.linear_layers:
        h = torch.relu(linear(h, lin.weight, lin.bias))
    h = h.reshape(-1, 1)
    for conv in net.out_convolutions:
        h = torch.relu(_conv_rows(h, conv, geom))
    return h


def _radial_log_prob(b, z: torch.Tensor) -> torch.Tensor:
    """Differentiable Lp-radial log-density (distributions.py:501-549) with the log-normal and (generalised) Gamma families
    of radius distributions."""
    v = z - b.loc.reshape(-1)
    r = v.abs().sum(-1) if b.p == 1.0 else v.pow(2).sum(-1).sqrt() if b.p == 2.0 else v.abs().max(-1).values
    logr = r.log()
    return radius_log_prob(b.norm_distribution, r, logr) - (b.log_delta_volume_const() + (b.dim - 1) * logr)


def radius_log_prob(nd, r: torch.Tensor, logr: torch.Tensor) -> torch.Tensor:
    """log f_R(r) of a radius distribution of `usflows_b200.distributions`, as a torch expression of its parameters."""
    if hasattr(nd, "_lognormals"):
        logits, mu, sg = nd._lognormals()
        t = torch.log_softmax(logits, 0) - ((logr[:, None] - mu) ** 2) / (2 * sg ** 2) - sg.log() - 0.5 * math.log(2 * math.pi)
        return torch.logsumexp(t, -1) - logr
    if hasattr(nd, "_mixture"):
        logits, a, rate, scale, power = nd._mixture()
        t = torch.log_softmax(logits, 0) + a * rate.log() - torch.lgamma(a)
        if scale is None:
            t = t + torch.xlogy(a - 1, r[:, None]) - rate * r[:, None]
        else:                                    # R = scale S^(1 / power): Chi (distributions.py:88-97), Weibull, HalfNormal
            lu = logr[:, None] - scale.log()
            t = t + power.log() - scale.log() + (a * power - 1) * lu - rate * torch.exp(power * lu)
        return torch.logsumexp(t, -1)
    raise NotImplementedError(f"usflows_b200: training with radius distribution {type(nd).__name__} is not built")


def base_log_prob(base, z: torch.Tensor) -> torch.Tensor:
    """Differentiable Laplace / Normal log-density summed over the event (distributions.py:150-151, 199-238)."""
    from .distributions import Independent, RadialDistribution
    b = base.base_dist if isinstance(base, Independent) else base
    if isinstance(b, RadialDistribution):
        return _radial_log_prob(b, z)
    loc = b.loc.reshape(-1)
    raw = b.scale_unconstrained
    scale = torch.nn.functional.softplus(raw.expand_as(b.loc) if raw.dim() == 0 else raw).reshape(-1)
    if b.base_kind == ops.BASE_LAPLACE:
        lp = -torch.log(2 * scale) - (z - loc).abs() / scale
    else:
        lp = -((z - loc) ** 2) / (2 * scale ** 2) - scale.log() - 0.5 * math.log(2 * math.pi)
    return lp.sum(-1)


def apply_autograd(flow, x: torch.Tensor, direction: str, context: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`Flow.backward` ("backward": data -> latent) / `Flow._forward` ("forward") on the autograd route -- the route that
    takes a context (flows.py:45-67 with the `context` of :235-238, 257-263)."""
    ev = tuple(flow._event_shape())
    for net in getattr(flow, "_cond_dense_nets", ()):      # `backward` / `_forward` without a context skip the context layer
        net.zero_context_default = False
    geom = ev if len(ev) == 3 else None
    if geom is not None:
        C, H, W = ev
        z = x.reshape(-1, C, H * W).transpose(1, 2).reshape(-1, C)
    else:
        z = x.reshape(-1, ev[0])
    cache: dict = {}
    for layer in (reversed(flow.layers) if direction == "backward" else flow.layers):
        z, _ = _layer_backward(layer, z, inverse=direction != "backward", cache=cache, geom=geom, context=context)
    if geom is not None:
        z = z.reshape(-1, H * W, C).transpose(1, 2).reshape(-1, C, H, W)
    return z


def log_prob_autograd(flow, x: torch.Tensor, context: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`Flow.log_prob` (flows.py:225-245) as an autograd graph over the flow's parameters; `context` [N] / [N, 1] reaches
    the conditional conditioners (soft training, flows.py:172-193)."""
    ev = tuple(flow._event_shape())
    for net in getattr(flow, "_cond_dense_nets", ()):      # log_prob without a context: the zero context of a soft USFlow
        net.zero_context_default = bool(getattr(flow, "_zero_ctx", False))
    geom = None
    if len(ev) == 3:                             # image-shaped event: channels-last rows [N*H*W, C] through the layers
        geom = ev
        C, H, W = ev
        z = x.reshape(-1, C, H * W).transpose(1, 2).reshape(-1, C)
    else:
        z = x.reshape(x.shape[0], -1)
    total = z.new_zeros(())                      # scalar, or [rows] once a data-dependent log-det joins
    cache: dict = {}
    for layer in reversed(flow.layers):
        z, ladj = _layer_backward(layer, z, cache=cache, geom=geom, context=context)
        total = total + ladj
    if geom is not None:                         # back to the NCHW element order the base parameters are stored in
        z = z.reshape(-1, H * W, C).transpose(1, 2).reshape(-1, C * H * W)
    return base_log_prob(flow.base_distribution, z) - total


# --------------------------------------------------------------------------------------------------
# data parallel plumbing
# --------------------------------------------------------------------------------------------------
def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, near-equal slice [lo, hi) of n rows for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(params: List[torch.nn.Parameter], group=None) -> int:
    """One blocking all-reduce(sum) over the flat fp32 gradient buffer; returns the number of floats exchanged.
    Parameters without a gradient on this rank (an empty shard) contribute zeros.  (`GradReducer` is the bucketed,
    overlapped form `fit` and `TrainStep` use; this one stays as the simple reference for the tests.)"""
    import torch.distributed as dist
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return off


class GradReducer:
    """Bucketed gradient all-reduce overlapped with the backward pass.

    Parameters are grouped into buckets in the order in which the density pass's backward produces their gradients: base
    density first, then the layers in registration order (the pass walks the layer list backwards, so its backward reaches
    coupling block 0 first and the final affine / scale layers last).  A post-accumulate-grad hook per parameter (it fires once per backward pass, after the contributions of a block and of
    the InverseTransform that aliases it have been summed) copies the gradient into its bucket; the bucket that just
    became complete goes out as ONE asynchronous all-reduce(sum) (NCCL on its own stream over NVLink / NVSwitch; gloo in
    the CPU tests) while the backward pass keeps running.  `finish()` waits for the buckets in flight, sends the ones a
    rank could not complete (parameters without a gradient contribute zeros, e.g. an empty shard) and copies the sums
    back into `p.grad`.  The result equals `allreduce_gradients` bit for bit on two ranks (sum of two numbers)."""

    def __init__(self, params, group=None, bucket_bytes: int = 16 << 20, first: Optional[list] = None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.params = [p for p in params if p.requires_grad]
        # expected order of readiness: `first` (the base density's parameters), then REGISTRATION order -- the density pass
        # walks the layer list backwards (flows.py:235), so its backward pass reaches block 0 first and the scale layer last
        head = [p for p in (first or []) if p.requires_grad]
        ids = {id(p) for p in head}
        order = head + [p for p in self.params if id(p) not in ids]
        self.buckets: List[dict] = []
        cur, cur_bytes = [], 0
        for p in order:
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self.buckets.append(dict(params=cur))
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(dict(params=cur))
        self._where = {}
        for bi, b in enumerate(self.buckets):
            n = sum(p.numel() for p in b["params"])
            ref = b["params"][0]
            b["flat"] = torch.zeros(n, dtype=torch.float32, device=ref.device)
            b["views"], off = [], 0
            for p in b["params"]:
                b["views"].append(b["flat"][off:off + p.numel()].view_as(p))
                self._where[id(p)] = (bi, len(b["views"]) - 1)
                off += p.numel()
            b["ready"], b["work"] = 0, None
        self.floats = sum(b["flat"].numel() for b in self.buckets)
        self.bytes_per_step = 4 * self.floats
        self._armed = False
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def close(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def begin(self) -> None:
        """Arm the hooks for one backward pass."""
        for b in self.buckets:
            b["ready"], b["work"] = 0, None
            b["filled"] = [False] * len(b["params"])
        self._armed = True

    def _launch(self, b) -> None:
        b["work"] = self.dist.all_reduce(b["flat"], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)

    def push(self, p) -> None:
        """Hand-written backward passes (train_engine.py) announce a finished gradient here (no autograd hook fires)."""
        self._on_grad(p)

    def _on_grad(self, p) -> None:
        if not self._armed:
            return
        bi, vi = self._where[id(p)]
        b = self.buckets[bi]
        if b["filled"][vi]:
            return
        if p.grad.data_ptr() != b["views"][vi].data_ptr():     # (the engine writes its gradients straight into the views)
            b["views"][vi].copy_(p.grad)
        b["filled"][vi] = True
        b["ready"] += 1
        if b["ready"] == len(b["params"]):
            self._launch(b)

    def finish(self) -> int:
        """Complete the exchange of this step; returns the number of floats all-reduced."""
        self._armed = False
        for b in self.buckets:
            if b["work"] is None:                       # incomplete on this rank: missing gradients are zeros
                for vi, p in enumerate(b["params"]):
                    if not b["filled"][vi]:
                        if p.grad is None:
                            b["views"][vi].zero_()
                        elif p.grad.data_ptr() != b["views"][vi].data_ptr():
                            b["views"][vi].copy_(p.grad)
                self._launch(b)
        for b in self.buckets:
            b["work"].wait()
            for p, v in zip(b["params"], b["views"]):
                if p.grad is None:
                    p.grad = v.clone()
                elif p.grad.data_ptr() != v.data_ptr():
                    p.grad.copy_(v)
        return self.floats

    def view_of(self, p) -> torch.Tensor:
        """The slice of the bucket buffer that holds `p`'s gradient: a backward pass that writes there directly (and sets
        `p.grad` to it) saves the copy into the bucket and the copy back."""
        bi, vi = self._where[id(p)]
        return self.buckets[bi]["views"][vi]


class TrainStep:
    """One maximum-likelihood step of `Flow.fit` (flows.py:195-207) as a reusable object: zero_grad, density pass and its
    backward on the library's kernels, the (overlapped) gradient all-reduce when a process group is given, optimiser
    step, and the invertibility check as a DEVICE counter (`infeasible` accumulates the number of zero diagonal / scale
    entries seen after each step; the caller reads it when it wants to, not once per step).  `global_rows` is the size of
    the global batch (the loss is the global mean, so shard gradients add up to the single-process gradient)."""

    def __init__(self, flow, opt, group=None, distributed: Optional[bool] = None, gradient_clip: Optional[float] = None,
                 engine: Optional[bool] = None, graph: Optional[bool] = None):
        import torch.distributed as dist
        from . import train_engine
        self.flow, self.opt, self.group, self.clip = flow, opt, group, gradient_clip
        # hand-written forward / backward (train_engine.py) where the flow is in its scope, torch autograd over the
        # library's contractions (`log_prob_autograd`) otherwise; `engine=False` forces the autograd route
        prior = flow.log_prior()
        self.use_engine = (engine is not False) and train_engine.supports(flow) \
            and not isinstance(prior, torch.Tensor) and prior == 0
        if engine is True and not self.use_engine:
            raise NotImplementedError("usflows_b200: this flow is outside the scope of the hand-written training pass")
        self._engines: Dict[int, Any] = {}
        # whole-step CUDA graphs: for the optimisers known to be capturable (this package's SophiaG; torch optimisers
        # built with capturable=True), unless switched off
        from .optim import SophiaG
        capturable = isinstance(opt, SophiaG) or all(g.get("capturable", False) for g in opt.param_groups)
        self.graph = bool(capturable) if graph is None else bool(graph)
        self._graphs: Dict[tuple, dict] = {}
        self._graph_warm: Dict[tuple, int] = {}
        self._lr_sig = self._lr_signature()
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized()
        self.distributed = distributed
        self.rank = dist.get_rank(group) if distributed else 0
        self.world = dist.get_world_size(group) if distributed else 1
        self.params = [p for p in flow.parameters()]
        base_params = list(flow.base_distribution.parameters())
        self.reducer = GradReducer(self.params, group, first=base_params) if distributed and self.world > 1 else None
        dev = self.params[0].device
        self.infeasible = torch.zeros((), dtype=torch.float32, device=dev)
        self.out_of_range = torch.zeros((), dtype=torch.float32, device=dev)   # fp16-split range flag of the engine passes
        self.allreduce_bytes = self.reducer.bytes_per_step if self.reducer is not None else 0

    def close(self) -> None:
        if self.reducer is not None:
            self.reducer.close()

    # -- CUDA-graph replay of the whole step ------------------------------------------------------------------------
    # The hand-written pass is ~400 launches of 5-60 us: issued one by one from Python (ctypes + tensor-map encoding,
    # ~15 us each) the host bounds the step.  Every shape and every buffer of `TrainEngine` is static, so after two
    # eager steps the sequence [engine pass -> gradient exchange -> clip -> optimiser -> invertibility check] is
    # captured once per (rows, global batch) and replayed; the NCCL all-reduces of the buckets are captured with it
    # (on NCCL's stream, so they still overlap the remaining backward kernels).  Optimisers that cannot be captured
    # (host-side state reads) keep the eager route: `graph=False`, or automatically when the capture fails.
    def step(self, sample: torch.Tensor, global_rows: Optional[int] = None,
             context: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Runs the step on this rank's shard `sample` [rows, ...]; returns this rank's share of the loss (a device
        scalar: the sum over ranks is the global mean loss).  `context` [rows, 1]: the noise scales of soft training."""
        rows = sample.shape[0]
        total = global_rows if global_rows is not None else rows * self.world
        if context is not None:
            return self._eager_step(sample, rows, total, context=context)
        if self.graph and self.use_engine and rows > 0 and sample.is_cuda and float(self._lr_signature()) == self._lr_sig:
            key = (rows, total)
            if key in self._graphs:
                ent = self._graphs[key]
                ent["x"].copy_(sample.reshape(rows, -1))
                ent["graph"].replay()
                return ent["loss"].clone()
            n = self._graph_warm.get(key, 0)
            if n >= 2:
                got = self._capture(sample, rows, total)
                if got is not None:
                    return got
            else:
                self._graph_warm[key] = n + 1
        return self._eager_step(sample, rows, total)

    def _lr_signature(self) -> float:
        """Hyper-parameters baked into a captured optimiser step (a scheduler changing them invalidates the graphs)."""
        return sum(float(g.get("lr", 0.0)) * (i + 1) for i, g in enumerate(self.opt.param_groups))

    def _capture(self, sample: torch.Tensor, rows: int, total: int):
        x_static = torch.empty(rows, sample.numel() // rows, dtype=torch.float32, device=sample.device)
        x_static.copy_(sample.reshape(rows, -1))
        out = {}
        graph = torch.cuda.CUDAGraph()
        try:
            from .flows import _plain_stream_order, capture_stream
            torch.cuda.synchronize(sample.device)
            with _plain_stream_order(), torch.cuda.graph(graph, stream=capture_stream(sample.device),
                                                         capture_error_mode="thread_local"):
                out["loss"] = self._eager_step(x_static, rows, total, zero_grad=False)
        except Exception as e:                                  # noqa: BLE001  (optimiser / collective not capturable)
            import warnings
            warnings.warn(f"usflows_b200: the training step is not CUDA-graph capturable here ({e!r}); running eagerly")
            self.graph = False
            torch.cuda.synchronize(sample.device)
            return None
        self._graphs[(rows, total)] = dict(graph=graph, x=x_static, loss=out["loss"])
        graph.replay()                                          # capturing does not execute: this is the step itself
        return out["loss"].clone()

    def _eager_step(self, sample: torch.Tensor, rows: int, total: int, zero_grad: bool = True,
                    context: Optional[torch.Tensor] = None) -> torch.Tensor:
        flow = self.flow
        if zero_grad:
            self.opt.zero_grad()
        if self.reducer is not None:
            self.reducer.begin()
        if rows > 0 and self.use_engine:
            from . import train_engine
            eng = self._engines.get(rows)
            if eng is None:
                if len(self._engines) >= 2:                    # full batches + one ragged last batch
                    self._engines.pop(next(iter(self._engines)))
                eng = self._engines[rows] = train_engine.TrainEngine(flow, rows)
                if self.reducer is not None:                   # gradients land in the all-reduce buckets directly
                    eng.bind_grad_buffers(self.reducer.view_of)
            eng.flag.zero_()
            local = eng.step(sample.reshape(rows, -1), total, self.reducer)
            self.out_of_range += eng.flag[0]
        elif rows > 0:
            loss = -log_prob_autograd(flow, sample, context).sum() / total
            if self.rank == 0:                          # the prior is a function of the weights only: count it once
                prior = flow.log_prior()
                if isinstance(prior, torch.Tensor) or prior != 0:
                    loss = loss - prior
            loss.backward()
            local = loss.detach()
        else:
            local = torch.zeros((), device=self.params[0].device)
        if self.reducer is not None:
            self.reducer.finish()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.clip)
        self.opt.step()
        self.infeasible += infeasible_count(flow)
        return local


def soft_training_noise(flow, batch: torch.Tensor, lo: int = 0, hi: Optional[int] = None):
    """SoftFlow perturbation of one GLOBAL batch (flows.py:172-193): one noise scale per sample from
    `training_noise_prior`, `x + N(0, sigma)` with that scale on every element of the sample, and the context
    `sigma * 2 / prior.high` the conditional conditioners see.  Drawn for the whole batch on every rank (same generator
    state => same draws) and cut to this rank's rows [lo, hi), so a data-parallel run perturbs the batch exactly as a
    single process would.  Returns (noisy rows, context [rows, 1])."""
    prior = flow.training_noise_prior
    n = batch.shape[0]
    hi = n if hi is None else hi
    sigma = prior.sample([n]).to(batch.device).reshape(n, *([1] * (batch.dim() - 1))).to(batch.dtype)
    noisy = batch + torch.randn_like(batch) * sigma
    context = (sigma.reshape(n, 1) * (2.0 / float(prior.high))).detach()
    return noisy[lo:hi], context[lo:hi]


def infeasible_count(flow) -> torch.Tensor:
    """Number of zero entries on the U diagonals / in the scale vectors of the flow's layers as a device scalar (the
    reference's `is_feasible`, transforms.py:150-152, 1347-1349, without the per-layer host synchronisation)."""
    from . import transforms as T
    tot = None
    seen = set()

    def visit(t):
        nonlocal tot
        if id(t) in seen:
            return
        seen.add(id(t))
        if isinstance(t, T.InverseTransform):
            visit(t.transform)
        elif isinstance(t, T.BlockAffineTransform):
            visit(t.block_transform)
        elif isinstance(t, T.SequentialAffineTransform):
            for s in t.transforms:
                visit(s)
        elif isinstance(t, T.LUTransform):
            c = (t.U_raw.detach().diagonal() == 0).sum()
            tot = c if tot is None else tot + c
        elif isinstance(t, T.ScaleTransform):
            c = (t.scale.detach() == 0).sum()
            tot = c if tot is None else tot + c

    for layer in flow.layers:
        visit(layer)
    if tot is None:
        return torch.zeros((), dtype=torch.float32, device=next(flow.parameters()).device)
    return tot.to(torch.float32)


def fit(flow, data_train, optim=None, optim_params: Optional[Dict[str, Any]] = None, batch_size: int = 32,
        shuffle: bool = True, gradient_clip: Optional[float] = None, device=None, epochs: int = 1,
        distributed: Optional[bool] = None, group=None, max_steps: Optional[int] = None,
        feasibility_every: int = 16) -> List[float]:
    """Reference-compatible `Flow.fit` (flows.py:113-210); `batch_size` is the GLOBAL batch when distributed.  The
    invertibility check of flows.py:204-205 runs on the device after every step and is READ every `feasibility_every`
    steps and at the end of every epoch (one host synchronisation per that many steps instead of one per layer per
    step); the losses stay on the device until the epoch ends."""
    import torch.distributed as dist
    from .optim import SophiaG
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else flow.device
    model = flow.to(device)
    optim = optim or SophiaG
    params = [p for p in model.parameters()]
    opt = optim(params, **optim_params) if optim_params is not None else optim(params)
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized()
    ts = TrainStep(model, opt, group=group, distributed=distributed, gradient_clip=gradient_clip)
    rank, world = ts.rank, ts.world

    def check_feasible():
        if float(ts.infeasible) != 0:                                                  # flows.py:204-205
            raise RuntimeError("Model is not invertible")
        if ts.use_engine and float(ts.out_of_range) != 0:
            # a value left the fp16-split range inside the hand-written pass (its gradients were not finite for the
            # affected steps): continue on the autograd route, whose tf32-split contractions have the fp32 range
            import warnings
            warnings.warn("usflows_b200: activations / gradients left the fp16-split range; training continues on the "
                          "tf32-split autograd route")
            ts.use_engine = False

    N = len(data_train)
    epoch_losses: List[float] = []
    steps = 0
    try:
        with ops.on_device(params[0]):
            for _ in range(epochs):
                losses = []
                perm = np.random.choice(N, N, replace=False) if shuffle else np.arange(N)     # flows.py:160
                if distributed:                                  # every rank walks the same permutation
                    pt = torch.from_numpy(perm).to(device if torch.device(device).type == "cuda" else "cpu")
                    dist.broadcast(pt, src=0, group=group)
                    perm = pt.cpu().numpy()
                data = data_train[perm] if isinstance(data_train, torch.Tensor) else data_train[perm][0]
                if not isinstance(data, torch.Tensor):
                    data = torch.as_tensor(np.asarray(data), dtype=torch.float32)
                for idx in range(0, N, batch_size):
                    end = min(idx + batch_size, N)
                    lo, hi = shard_bounds(end - idx, rank, world)
                    sample = data[idx + lo:idx + hi].to(device=device, dtype=torch.float32)
                    context = None
                    if flow.soft_training:                       # flows.py:172-193
                        sample, context = soft_training_noise(flow, data[idx:end].to(device=device, dtype=torch.float32), lo, hi)
                    losses.append(ts.step(sample, global_rows=end - idx, context=context))
                    steps += 1
                    if steps % max(1, feasibility_every) == 0:
                        check_feasible()
                    if max_steps is not None and steps >= max_steps:
                        break
                check_feasible()
                if losses:
                    tot = torch.stack(losses)
                    if distributed and world > 1:               # one exchange per epoch: the per-step global losses
                        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
                    epoch_losses.append(float(tot.mean()))
                else:
                    epoch_losses.append(float("nan"))
                if max_steps is not None and steps >= max_steps:
                    break
    finally:
        ts.close()
    return epoch_losses
