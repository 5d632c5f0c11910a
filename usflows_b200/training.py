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


def _layer_backward(layer, y: torch.Tensor, inverse: bool = False, cache: Optional[dict] = None, geom=None):
    """(density direction value, forward log|det J|) of one layer; `inverse` swaps the direction.  For image-shaped events
    `geom = (C, H, W)` and y holds channels-last rows [N*H*W, C] (the layout of image_engine.py): the 1x1 convolution of a
    BlockAffineTransform is then the same `linear` as the flat case, per-element vectors are indexed per pixel."""
    from . import transforms as T
    if isinstance(layer, T.InverseTransform):
        x, ladj = _layer_backward(layer.transform, y, not inverse, cache, geom)
        return x, -ladj
    if isinstance(layer, T.BlockAffineTransform):
        W, Winv, b, ladj = affine_parts(layer.block_transform, cache)
        ladj = ladj * layer.n_blocks
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
        t = _convnet2d_rows(layer.conditioner, (y3 * m).reshape(-1, C), geom).reshape(-1, H * W, C)
        t = (1 - m) * t
        return ((y3 + t) if inverse else (y3 - t)).reshape(-1, C), y.new_zeros(())
    if isinstance(layer, T.MaskedCoupling):
        m = layer.mask.reshape(-1).to(y.dtype)
        t = (1 - m) * _conditioner(layer.conditioner, y * m)
        return (y + t if inverse else y - t), y.new_zeros(())      # transforms.py:277-306, 316-326
    raise NotImplementedError(f"usflows_b200: training of {type(layer).__name__} is not built")


def _conditioner(net, h: torch.Tensor) -> torch.Tensor:
    """Conditioner network as an autograd graph: contractions on the library's kernels, element-wise glue in torch."""
    from .nn import ConvNet
    if isinstance(net, ConvNet):                      # networks.py:222-245 (GatedMLP), 205-219 (LayerNormVector), 287-307
        d = net._describe()
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


def _convnet2d_rows(net, x: torch.Tensor, geom) -> torch.Tensor:
    """networks.ConvNet2D (networks.py:441-506) over channels-last rows as an autograd graph."""
    d = net._describe()
    h = _conv_rows(x, d["first"], geom)
    for blk in d["blocks"]:
        if blk["gated"]:                                  # GatedConv.forward, networks.py:103-121
            o = _conv_rows(torch.relu(_conv_rows(torch.relu(h), blk["conv1"], geom)), blk["conv2"], geom)
            val, gate = o.chunk(2, dim=1)
            h = h + val * torch.sigmoid(gate)
        else:
            h = _conv_rows(h, blk["conv1"], geom)
        h = torch.relu(h)
        if blk["ln"] is not None:                         # LayerNormChannels, networks.py:53-58 (per pixel = per row)
            ln = blk["ln"]
            h = torch.nn.functional.layer_norm(h, (h.shape[1],), ln.gamma.reshape(-1), ln.beta.reshape(-1), ln.eps)
    return _conv_rows(h, d["last"], geom)


def _radial_log_prob(b, z: torch.Tensor) -> torch.Tensor:
    """Differentiable Lp-radial log-density (distributions.py:501-549) with LogNormal / GammaMM radius distributions."""
    from .distributions import GammaMM, LogNormal
    v = z - b.loc.reshape(-1)
    r = v.abs().sum(-1) if b.p == 1.0 else v.pow(2).sum(-1).sqrt() if b.p == 2.0 else v.abs().max(-1).values
    logr = r.log()
    nd = b.norm_distribution
    sp = torch.nn.functional.softplus
    if isinstance(nd, LogNormal):
        mu, sg = nd.loc.reshape(()), sp(nd.scale_unconstrained).reshape(())
        lp = -((logr - mu) ** 2) / (2 * sg ** 2) - sg.log() - 0.5 * math.log(2 * math.pi) - logr
    elif isinstance(nd, GammaMM):
        a, rate = sp(nd.concentration_unconstrained), sp(nd.rate_unconstrained)
        t = torch.log_softmax(nd.mixture_logits, 0) + a * rate.log() - torch.lgamma(a) \
            + torch.xlogy(a - 1, r[:, None]) - rate * r[:, None]
        lp = torch.logsumexp(t, -1)
    else:
        raise NotImplementedError(f"usflows_b200: training with radius distribution {type(nd).__name__} is not built")
    return lp - (b.log_delta_volume_const() + (b.dim - 1) * logr)


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


def log_prob_autograd(flow, x: torch.Tensor) -> torch.Tensor:
    """`Flow.log_prob` (flows.py:225-245) as an autograd graph over the flow's parameters."""
    ev = tuple(flow._event_shape())
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
        z, ladj = _layer_backward(layer, z, cache=cache, geom=geom)
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
    """One all-reduce(sum) over the flat fp32 gradient buffer; returns the number of floats exchanged.
    Parameters without a gradient on this rank (an empty shard) contribute zeros."""
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


def fit(flow, data_train, optim=None, optim_params: Optional[Dict[str, Any]] = None, batch_size: int = 32,
        shuffle: bool = True, gradient_clip: Optional[float] = None, device=None, epochs: int = 1,
        distributed: Optional[bool] = None, group=None, max_steps: Optional[int] = None) -> List[float]:
    """Reference-compatible `Flow.fit` (flows.py:113-210); `batch_size` is the GLOBAL batch when distributed."""
    import torch.distributed as dist
    from .optim import SophiaG
    if flow.soft_training:
        raise NotImplementedError("usflows_b200: soft_training is not built")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else flow.device
    model = flow.to(device)
    optim = optim or SophiaG
    params = [p for p in model.parameters()]
    opt = optim(params, **optim_params) if optim_params is not None else optim(params)
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1

    N = len(data_train)
    epoch_losses: List[float] = []
    steps = 0
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
            opt.zero_grad()
            if hi > lo:
                loss = -log_prob_autograd(model, sample).sum() / (end - idx) - model.log_prior()
                loss.backward()
                local = loss.detach()
            else:
                local = torch.zeros((), device=device)
            if distributed:
                allreduce_gradients(params, group)
                dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
            losses.append(float(local))
            if gradient_clip is not None:
                torch.nn.utils.clip_grad_norm_(params, gradient_clip)
            opt.step()
            if not model.is_feasible():                                                # flows.py:204-205
                raise RuntimeError("Model is not invertible")
            steps += 1
            if max_steps is not None and steps >= max_steps:
                break
        epoch_losses.append(float(np.mean(losses)) if losses else float("nan"))
        if max_steps is not None and steps >= max_steps:
            break
    return epoch_losses
