"""Conditioner network of the coupling layers: drop-in for `pyro.nn.DenseNN` (pyro-ppl 1.8.6), which the
reference imports for its flat/MLP configurations (experiments/synthetic/gaussian_mixture.yaml:66-71).

Only a parameter container: `Linear(d,H0) -> ReLU -> ... -> Linear(Hk, sum(param_dims))`, same attribute
and state-dict names (`layers.{j}.weight|bias`).  The arithmetic runs fused inside
`MaskedCoupling` (tcgen05 / SIMT contraction kernels with bias+ReLU epilogues); calling the module directly
evaluates the MLP through the same kernels.
"""
from __future__ import annotations

import torch


class DenseNN(torch.nn.Module):
    def __init__(self, input_dim, hidden_dims, param_dims=[1, 1], nonlinearity=torch.nn.ReLU()):
        super().__init__()
        if not isinstance(nonlinearity, torch.nn.ReLU):
            raise NotImplementedError("usflows_b200.nn.DenseNN: only the ReLU nonlinearity is fused")
        self.input_dim = input_dim
        self.hidden_dims = list(hidden_dims)
        self.param_dims = list(param_dims)
        self.count_params = len(param_dims)
        self.output_multiplier = sum(param_dims)
        dims = [input_dim] + self.hidden_dims + [self.output_multiplier]
        self.layers = torch.nn.ModuleList(
            [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])
        self.f = nonlinearity

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import engine
        out = engine.run_mlp(self, x)
        if self.count_params == 1:
            return out
        return tuple(out.split(self.param_dims, dim=-1))       # pyro.nn.DenseNN returns one tensor per entry of param_dims


class AdditiveAffineNN(torch.nn.Module):
    """`networks.AdditiveAffineNN` (networks.py:14-38): a DenseNN `loc_fnc` that yields `[loc, log_scale]` with
    `log_scale = 0` -- the parameters of a purely additive affine transform (state-dict keys `loc_fnc.layers.<i>. ...`)."""

    def __init__(self, input_dim: int, hidden_dims, output_dim: int, nonlinearity=None):
        super().__init__()
        self.loc_fnc = DenseNN(input_dim, hidden_dims, [output_dim],
                               nonlinearity=torch.nn.ReLU() if nonlinearity is None else nonlinearity)

    def forward(self, x: torch.Tensor):
        loc = self.loc_fnc(x)
        return [loc, torch.zeros_like(loc)]


class _DenseView:
    """weight / bias of one Linear as the contraction kernels will see it (`bias` may be a sum of two parameters)."""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class _MLPView:
    def __init__(self, layers):
        self.layers = layers


def mlp_layers(net) -> list:
    """The Linear / ReLU stack a coupling lowers for a `DenseNN` or a `ConditionalDenseNN` evaluated without an explicit
    context (objects with `.weight` [out, in] and `.bias`)."""
    return net._mlp_layers() if hasattr(net, "_mlp_layers") else list(net.layers)


class ConditionalDenseNN(torch.nn.Module):
    """`networks.ConditionalDenseNN` (networks.py:681-752): a Linear / ReLU stack whose first layer adds a Linear of the
    context, `h = f(L0 x + L1 c)`; `layers = [L0 (input), L1 (context), hidden ..., output]` with the reference's state-dict
    keys (`layers.<i>.weight / bias`).  The soft-training conditioner for flat events (flows.py:172-193 hands every
    coupling the per-sample noise scale as `context` [N, 1]).

    Without an explicit context the stack is a plain MLP and runs on the fused coupling launches: `context=None` leaves
    `L1` out altogether (networks.py:739-741), the zero context a soft-training `USFlow` substitutes (flows.py:559-565)
    contributes `L1`'s bias, which is folded into `L0`'s (`zero_context_default`, set by the flow).  With a context the
    first layer gets the rank-`context_dim` term on the layer-by-layer route (training.py)."""

    def __init__(self, input_dim, context_dim, hidden_dims, out_dim, nonlinearity=torch.nn.ReLU()):
        super().__init__()
        _require_relu(nonlinearity)
        self.input_dim, self.context_dim, self.hidden_dims, self.out_dim = input_dim, context_dim, list(hidden_dims), out_dim
        layers = [torch.nn.Linear(input_dim, self.hidden_dims[0]), torch.nn.Linear(context_dim, self.hidden_dims[0])]
        for i in range(1, len(self.hidden_dims)):
            layers.append(torch.nn.Linear(self.hidden_dims[i - 1], self.hidden_dims[i]))
        layers.append(torch.nn.Linear(self.hidden_dims[-1], out_dim))
        self.layers = torch.nn.ModuleList(layers)
        self.f = nonlinearity
        self.zero_context_default = False

    @property
    def context_channels(self) -> int:
        return self.context_dim

    def _mlp_layers(self, zero_context=None) -> list:
        zero = self.zero_context_default if zero_context is None else zero_context
        first = self.layers[0]
        if zero:
            first = _DenseView(first.weight, first.bias + self.layers[1].bias)
        return [first] + list(self.layers[2:])

    def forward(self, x: torch.Tensor, context=None) -> torch.Tensor:
        if context is None:
            from . import engine
            return engine.run_mlp(_MLPView(self._mlp_layers(False)), x)
        from . import training
        with torch.no_grad():
            return training._conditioner(self, x.reshape(-1, x.shape[-1]), context).reshape(*x.shape[:-1], self.out_dim)


# --------------------------------------------------------------------------------------------------
# The reference's own MLP-style conditioner: `networks.ConvNet` with 1-D in_dims (networks.py:205-245, 248-307, 379-389)
# --------------------------------------------------------------------------------------------------
def _require_relu(nonlinearity) -> None:
    if not isinstance(nonlinearity, torch.nn.ReLU):
        raise NotImplementedError("usflows_b200.nn: only the ReLU nonlinearity is fused")


class LayerNormVector(torch.nn.Module):
    """Parameter container of a LayerNorm over the feature dim (networks.py:205-219; keys `layernorm.weight|bias`)."""

    def __init__(self, features: int, eps: float = 1e-5):
        super().__init__()
        self.layernorm = torch.nn.LayerNorm(features, eps=eps)


class GatedMLP(torch.nn.Module):
    """x [-> proj] + val * sigmoid(gate), [val | gate] = Linear(out, 2 out)(relu(Linear(in, out)(relu(x))))
    (networks.py:222-245).  Parameter container: keys `net1.1.*`, `net1.3.*`, `proj.*` as in the reference."""

    def __init__(self, in_features: int, out_features: int, nonlinearity=torch.nn.ReLU()):
        super().__init__()
        _require_relu(nonlinearity)
        self.net1 = torch.nn.Sequential(nonlinearity, torch.nn.Linear(in_features, out_features), nonlinearity,
                                        torch.nn.Linear(out_features, 2 * out_features))
        self.proj = torch.nn.Linear(in_features, out_features) if in_features != out_features else None


class ConvNet(torch.nn.Module):
    """`networks.ConvNet` for vector inputs (`in_dims=[d]`): Linear(d, h0) -> [GatedMLP | ReLU+Linear -> LayerNormVector]*
    -> Linear(h_last, c_out), same constructor arguments and state-dict names (`nn.{i}. ...`) as the reference
    (networks.py:262-307).  A parameter container: inside a `MaskedCoupling` its contractions run on the tcgen05 /
    SIMT kernels and the gate / LayerNorm glue on `usf_gate_norm`; calling the module evaluates it through the same
    kernels.  `in_dims=[C, H, W]` builds the convolutional branch (networks.py:308-377; 2-D only), which runs like a
    `ConvNet2D` (image_engine.py) with per-block widths and residual projections."""

    def __init__(self, in_dims, c_hidden, c_out: int = -1, nonlinearity=torch.nn.ReLU(), kernel_size: int = 3,
                 stride: int = 1, dilation: int = 1, padding=None, normalize_layers: bool = True, gating: bool = True):
        super().__init__()
        _require_relu(nonlinearity)
        try:
            in_dims = list(in_dims)
        except TypeError:
            raise ValueError("in_dims must be an iterable like [C, H, W] or [C] for vector")
        c_in = int(in_dims[0])
        c_out = c_out if c_out > 0 else c_in
        assert len(c_hidden) > 0 and all(h > 0 for h in c_hidden), "c_hidden must be non-empty list of positive ints"
        hidden = [int(h) for h in c_hidden]
        if len(in_dims) != 1:
            self._build_spatial(in_dims, c_in, hidden, c_out, nonlinearity, kernel_size, stride, dilation, padding,
                                normalize_layers, gating)
            return
        layers = [torch.nn.Linear(c_in, hidden[0])]
        for i, out_ch in enumerate(hidden):
            in_ch = hidden[i - 1] if i > 0 else hidden[0]
            if gating:
                layers.append(GatedMLP(in_ch, out_ch, nonlinearity=nonlinearity))
            else:
                layers.append(torch.nn.Sequential(nonlinearity, torch.nn.Linear(in_ch, out_ch)))
            if normalize_layers:
                layers.append(LayerNormVector(out_ch))
        layers.append(torch.nn.Linear(hidden[-1], c_out))
        self.nn = torch.nn.Sequential(*layers)
        self.is_vector = True
        self._vector_in_features = c_in

    def _build_spatial(self, in_dims, c_in, hidden, c_out, nonlinearity, kernel_size, stride, dilation, padding,
                       normalize_layers, gating) -> None:
        """The convolutional branch (networks.py:308-377) for `in_dims=[C, H, W]`: Conv k x k (C, h0) -> [GatedConvND |
        Conv k x k -> ReLU -> LayerNormChannelsND]* -> Conv k x k (h_last, c_out); a block whose width changes projects
        its residual with a 1x1 convolution (networks.py:186-201).  Same state-dict names as the reference."""
        if len(in_dims) != 3:
            raise NotImplementedError("usflows_b200.nn.ConvNet: in_dims=[d] (vector) and [C, H, W] (2-D convolutions) are built")
        if padding is None:
            padding = kernel_size // 2
        if stride != 1 or kernel_size % 2 != 1 or not (padding == "same" or padding == (kernel_size // 2) * dilation):
            raise NotImplementedError("usflows_b200.nn.ConvNet: stride 1, odd kernel_size and 'same'-sized padding only")
        conv = lambda i, o: torch.nn.Conv2d(i, o, kernel_size=kernel_size, padding=padding, stride=stride,   # noqa: E731
                                            dilation=dilation)
        layers = [conv(c_in, hidden[0])]
        for i, out_ch in enumerate(hidden):
            in_ch = hidden[i - 1] if i > 0 else hidden[0]
            if gating:
                layers += [GatedConvND(in_ch, out_ch, kernel_size=kernel_size, padding=padding, stride=stride,
                                       dilation=dilation, nonlinearity=nonlinearity, input_rank=2), nonlinearity]
            else:
                layers += [conv(in_ch, out_ch), nonlinearity]
            if normalize_layers:
                layers += [LayerNormChannelsND(out_ch, num_spatial_dims=2)]
        layers += [conv(hidden[-1], c_out)]
        self.nn = torch.nn.Sequential(*layers)
        self.is_vector = False
        self._spatial_rank = 2
        self.kernel_size, self.dilation = kernel_size, dilation

    context_channels = 0          # CondConvNet: trailing input channels / features that carry the context

    def _describe(self, with_context: bool = False) -> dict:
        """Structure of the network as the engine's planner reads it: first / last Linear and the hidden blocks (the
        spatial branch answers in `ConvNet2D._describe`'s format, with `proj` per block).  A conditional network
        (`CondConvNet`) is described WITHOUT its context inputs unless `with_context`: context 0, which is what the
        reference evaluates `log_prob` / `backward` / `sample` with when none is given (flows.py:559-565)."""
        if not self.is_vector:
            return _describe_conv_stack(self.nn, self.kernel_size, self.dilation, 0 if with_context else self.context_channels)
        mods = list(self.nn)
        blocks, i = [], 1
        while i < len(mods) - 1:
            m = mods[i]
            if isinstance(m, GatedMLP):
                blk = dict(gated=True, lin1=m.net1[1], lin2=m.net1[3], proj=m.proj, ln=None)
            else:
                blk = dict(gated=False, lin1=m[1], lin2=None, proj=None, ln=None)
            i += 1
            if i < len(mods) - 1 and isinstance(mods[i], LayerNormVector):
                blk["ln"] = mods[i].layernorm
                i += 1
            blocks.append(blk)
        first = mods[0] if with_context or not self.context_channels else _without_context(mods[0], self.context_channels)
        return dict(first=first, blocks=blocks, last=mods[-1])

    def forward(self, x: torch.Tensor, context=None) -> torch.Tensor:
        if context is not None and self.context_channels:
            raise NotImplementedError("usflows_b200.nn: a conditional network is evaluated with a context inside a flow "
                                      "(Flow.log_prob(x, context) / fit); called directly it uses context 0")
        if not self.is_vector:
            from . import image_engine
            return image_engine.run_convnet2d(self, x)
        from . import engine
        if x.dim() == 3 and x.shape[-1] == 1:
            x = x.reshape(x.shape[0], x.shape[1])
        return engine.run_conditioner(self, x)


# --------------------------------------------------------------------------------------------------
# Convolutional conditioner of the image-shaped flows: `networks.ConvNet2D` (networks.py:40-121, 405-510)
# --------------------------------------------------------------------------------------------------
class LayerNormChannels(torch.nn.Module):
    """Parameter container of the per-pixel LayerNorm over channels (networks.py:40-58; keys `gamma`, `beta`)."""

    def __init__(self, c_in: int, eps: float = 1e-5):
        super().__init__()
        self.gamma = torch.nn.Parameter(torch.ones(1, c_in, 1, 1))
        self.beta = torch.nn.Parameter(torch.zeros(1, c_in, 1, 1))
        self.eps = eps


class GatedConv(torch.nn.Module):
    """x + val * sigmoid(gate), [val | gate] = Conv1x1(c_hidden, 2 c_in)(relu(Conv k x k(c_in, c_hidden)(relu(x))))
    (networks.py:61-121).  Parameter container: keys `net.1.*`, `net.3.*`."""

    def __init__(self, c_in: int, c_hidden: int, kernel_size: int = 3, padding=1, stride: int = 1,
                 nonlinearity=torch.nn.ReLU(), dilation: int = 1):
        super().__init__()
        _require_relu(nonlinearity)
        assert stride == 1, "Stride > 1 cannot be used to skip connection."
        self.net = torch.nn.Sequential(
            nonlinearity, torch.nn.Conv2d(c_in, c_hidden, kernel_size=kernel_size, padding=padding, stride=stride,
                                          dilation=dilation),
            nonlinearity, torch.nn.Conv2d(c_hidden, 2 * c_in, kernel_size=1, padding=0))


class LayerNormChannelsND(torch.nn.Module):
    """Parameter container of the channel LayerNorm of `networks.ConvNet`'s spatial branch (networks.py:122-140): `gamma`,
    `beta` of shape (1, C, 1, ..., 1)."""

    def __init__(self, c_in: int, num_spatial_dims: int = 2, eps: float = 1e-5):
        super().__init__()
        shape = (1, c_in) + (1,) * num_spatial_dims
        self.gamma = torch.nn.Parameter(torch.ones(*shape))
        self.beta = torch.nn.Parameter(torch.zeros(*shape))
        self.eps = eps


class GatedConvND(torch.nn.Module):
    """proj(x) + val * sigmoid(gate), [val | gate] = Conv1x1(c_out, 2 c_out)(relu(Conv k x k(c_in, c_out)(relu(x)))), `proj` a
    1x1 convolution when c_in != c_out, else the identity (networks.py:142-203).  Parameter container: keys `net.1.*`,
    `net.3.*`, `proj.*`.  2-D only here."""

    def __init__(self, c_in: int, c_out: int, kernel_size: int = 3, padding=1, stride: int = 1, dilation: int = 1,
                 nonlinearity=torch.nn.ReLU(), input_rank: int = 2):
        super().__init__()
        _require_relu(nonlinearity)
        assert stride == 1, "Stride > 1 cannot be used to skip connection."
        if input_rank != 2:
            raise NotImplementedError("usflows_b200.nn.GatedConvND: 2-D convolutions only")
        self.net = torch.nn.Sequential(
            nonlinearity, torch.nn.Conv2d(c_in, c_out, kernel_size=kernel_size, padding=padding, stride=stride,
                                          dilation=dilation),
            nonlinearity, torch.nn.Conv2d(c_out, 2 * c_out, kernel_size=1, padding=0, stride=1))
        self.proj = torch.nn.Conv2d(c_in, c_out, kernel_size=1, padding=0) if c_in != c_out else None


class _LayerView:
    """A Linear / Conv2d seen without its trailing `n` context inputs: a zero context contributes nothing, so the layer
    is the one with `weight[:, :-n]` (a differentiable view of the parameter)."""

    def __init__(self, layer, n: int):
        self.weight = layer.weight[:, :layer.weight.shape[1] - n].contiguous()
        self.bias = layer.bias
        self.dilation = getattr(layer, "dilation", (1, 1))


def _without_context(layer, n: int):
    return _LayerView(layer, n)


def _describe_conv_stack(seq, kernel_size: int, dilation: int, drop_context: int = 0) -> dict:
    """first / blocks / last of a convolutional conditioner stack (`ConvNet2D.nn`, the spatial `ConvNet.nn`);
    `drop_context` trailing input channels of the first convolution are left out (context 0)."""
    mods = [m for m in seq if not isinstance(m, torch.nn.ReLU)]
    blocks, i = [], 1
    while i < len(mods) - 1:
        m = mods[i]
        if isinstance(m, (GatedConv, GatedConvND)):
            blk = dict(gated=True, conv1=m.net[1], conv2=m.net[3], proj=getattr(m, "proj", None), ln=None)
        else:
            blk = dict(gated=False, conv1=m, conv2=None, proj=None, ln=None)
        i += 1
        if i < len(mods) - 1 and isinstance(mods[i], (LayerNormChannels, LayerNormChannelsND)):
            blk["ln"] = mods[i]
            i += 1
        blocks.append(blk)
    first = _without_context(mods[0], drop_context) if drop_context else mods[0]
    return dict(first=first, blocks=blocks, last=mods[-1], k=kernel_size, dilation=dilation)


class ConvNet2D(torch.nn.Module):
    """`networks.ConvNet2D`: Conv k x k (c_in, c_hidden) -> [GatedConv | Conv k x k, ReLU, LayerNormChannels] x num_layers
    -> Conv k x k (c_hidden, c_out), same constructor arguments and state-dict names as the reference
    (networks.py:405-494).  A parameter container: inside a `MaskedCoupling` over image-shaped `in_dims` every
    convolution runs as a row gather (usf_im2col) + contraction over channels-last rows, the gate / ReLU / LayerNorm
    glue on usf_gate_norm.  Shape-preserving convolutions only: stride 1 and padding 'same' (or kernel_size // 2 *
    dilation), which is what the reference's couplings need (transforms.py:284-290 adds the output to x)."""

    def __init__(self, c_in: int, c_hidden: int = 3, c_out: int = -1, num_layers: int = 3, nonlinearity=torch.nn.ReLU(),
                 kernel_size: int = 3, stride: int = 1, dilation: int = 1, padding=0, normalize_layers: bool = True,
                 gating: bool = True):
        super().__init__()
        _require_relu(nonlinearity)
        if padding is None:
            padding = kernel_size // 2
        if stride != 1 or kernel_size % 2 != 1 or not (padding == "same" or padding == (kernel_size // 2) * dilation):
            raise NotImplementedError("usflows_b200.nn.ConvNet2D: stride 1, odd kernel_size and padding 'same' only")
        self.nonlinearity = nonlinearity
        c_out = c_out if c_out > 0 else c_in
        conv = lambda i, o: torch.nn.Conv2d(i, o, kernel_size=kernel_size, padding=padding, stride=stride,   # noqa: E731
                                            dilation=dilation)
        layers = [conv(c_in, c_hidden)]
        for _ in range(num_layers):
            if gating:
                layers += [GatedConv(c_hidden, c_hidden, kernel_size=kernel_size, padding=padding, stride=stride,
                                     dilation=dilation), nonlinearity]
            else:
                layers += [conv(c_hidden, c_hidden), nonlinearity]
            if normalize_layers:
                layers += [LayerNormChannels(c_hidden)]
        layers += [conv(c_hidden, c_out)]
        self.nn = torch.nn.Sequential(*layers)
        self.kernel_size, self.dilation = kernel_size, dilation

    context_channels = 0          # CondConvNet2D: 1

    def _describe(self, with_context: bool = False) -> dict:
        return _describe_conv_stack(self.nn, self.kernel_size, self.dilation, 0 if with_context else self.context_channels)

    def forward(self, x: torch.Tensor, context=None) -> torch.Tensor:
        if context is not None and self.context_channels:
            raise NotImplementedError("usflows_b200.nn: a conditional network is evaluated with a context inside a flow "
                                      "(Flow.log_prob(x, context) / fit); called directly it uses context 0")
        from . import image_engine
        return image_engine.run_convnet2d(self, x)


# --------------------------------------------------------------------------------------------------
# Context-conditioned conditioners of soft training (flows.py:172-193, 559-565): `networks.CondConvNet` (networks.py:513-
# 600) and `networks.CondConvNet2D` (networks.py:603-680) append the per-sample context (the noise scale) as ONE extra
# input channel / feature, constant over the pixels of a sample.
# --------------------------------------------------------------------------------------------------
class CondConvNet(ConvNet):
    """`networks.CondConvNet`: a `ConvNet` over `[in_dims[0] + 1, ...]` inputs.  As in the reference `c_out` defaults to
    the widened input (pass `c_out=in_dims[0]` for a coupling)."""

    context_channels = 1

    def __init__(self, in_dims, c_hidden, c_out: int = -1, nonlinearity=torch.nn.ReLU(), kernel_size: int = 3,
                 stride: int = 1, dilation: int = 1, padding=None, normalize_layers: bool = True, gating: bool = True,
                 **kwargs):
        try:
            in_dims = list(in_dims)
        except TypeError:
            raise ValueError("in_dims must be an iterable like [C, H, W]")
        super().__init__(in_dims=[in_dims[0] + 1] + in_dims[1:], c_hidden=c_hidden, c_out=c_out, nonlinearity=nonlinearity,
                         kernel_size=kernel_size, stride=stride, dilation=dilation, padding=padding,
                         normalize_layers=normalize_layers, gating=gating, **kwargs)
        self._orig_in_dims = in_dims


class CondConvNet2D(ConvNet2D):
    """`networks.CondConvNet2D`: a `ConvNet2D` over `c_in + 1` input channels with `c_out = c_in` by default."""

    context_channels = 1

    def __init__(self, c_in: int, c_hidden: int = 3, c_out: int = -1, num_layers: int = 3, nonlinearity=torch.nn.ReLU(),
                 kernel_size: int = 3, stride: int = 1, dilation: int = 1, padding=None, **kwargs):
        if c_out < 0:
            c_out = c_in
        super().__init__(c_in=c_in + 1, c_hidden=c_hidden, c_out=c_out, num_layers=num_layers, nonlinearity=nonlinearity,
                         kernel_size=kernel_size, stride=stride, dilation=dilation, padding=padding, **kwargs)


class BottleneckConv(torch.nn.Module):
    """`networks.BottleneckConv` (networks.py:754-824): two 'same' convolutions down to ONE channel, two Linear layers over
    the flattened pixels, two convolutions back up to `c_in` channels, a ReLU after every layer (the output included).
    State-dict keys as the reference (`in_convolutions.<i>`, `linear_layers.<i>`, `out_convolutions.<i>`); `c_hidden_in` /
    `c_hidden_out` are accepted and unused, as there.  No configuration of the reference uses it; a flow with this
    conditioner is evaluated layer by layer (training.py: every convolution / Linear a contraction on the library's
    kernels), not through a fused launch program."""

    layer_route_only = True

    def __init__(self, c_in, c_hidden_in, c_hidden_out, in_dims, c_hidden: int = 3, nonlinearity=torch.nn.ReLU(),
                 kernel_size: int = 3):
        super().__init__()
        _require_relu(nonlinearity)
        self.in_dims = list(in_dims)
        self.n_pixels = 1
        for n in self.in_dims[1:]:
            self.n_pixels *= int(n)
        conv = lambda i, o: torch.nn.Conv2d(i, o, kernel_size=kernel_size, padding="same")   # noqa: E731
        self.in_convolutions = torch.nn.ModuleList([conv(c_in, c_hidden), conv(c_hidden, 1)])
        self.linear_layers = torch.nn.ModuleList([torch.nn.Linear(self.n_pixels, self.n_pixels) for _ in range(2)])
        self.out_convolutions = torch.nn.ModuleList([conv(1, c_hidden), conv(c_hidden, c_in)])
        self.nonlinearity = nonlinearity

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import training
        C, H, W = x.shape[-3:]
        rows = x.reshape(-1, C, H * W).transpose(1, 2).reshape(-1, C)              # channels-last rows
        with torch.no_grad():
            out = training._bottleneck_rows(self, rows, (C, H, W))
        return out.reshape(-1, H * W, out.shape[1]).transpose(1, 2).reshape(*x.shape[:-3], out.shape[1], H, W)
