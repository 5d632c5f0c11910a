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
