"""Base distributions as modules with constrained learnable parameters -- drop-in mirrors of the reference's
`DistributionModule`, `Laplace`, `Normal`, `Independent` (src/usflows/distributions.py:117-238, 709-728).

Parameter names match the reference (`loc`, `scale_unconstrained`, scale = softplus(scale_unconstrained)).
`log_prob` and `sample` run the fused base-density / Philox sampling kernels; no torch.distributions object is
rebuilt per call (the reference does that on every access, distributions.py:129-138).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
from torch.nn import Module, Parameter

from . import ops
from .utils import inv_softplus


class DistributionModule(Module):
    base_kind: int = -1

    def __init__(self, n_batch_dims: int = 0):
        super().__init__()
        self.n_batch_dims = n_batch_dims
        self._seed_offset = 0

    # -- prepared parameters (scale = softplus(scale_unconstrained)), cached per weight version ------
    def _prepared(self):
        key = tuple((p.data_ptr(), p._version) for p in (self.loc, self.scale_unconstrained))
        if getattr(self, "_prep_key", None) != key:
            ops.require_cuda(self.loc, "base_distribution.loc")
            with torch.no_grad():
                raw = self.scale_unconstrained.detach()
                if raw.dim() == 0:                                 # scalar scale expands to loc's shape (:228-232)
                    raw = raw.expand_as(self.loc)
                raw = raw.reshape(-1).contiguous()
                scale = torch.empty_like(raw)
                ops.softplus(raw, scale)
                loc = self.loc.detach().reshape(-1).contiguous()
            self._prep_cache, self._prep_key = (loc, scale), key
        return self._prep_cache

    @property
    def event_shape(self) -> torch.Size:
        return torch.Size(self.loc.shape[self.n_batch_dims:])

    @property
    def batch_shape(self) -> torch.Size:
        return torch.Size(self.loc.shape[:self.n_batch_dims])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.log_prob(x)

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        """sum over the event dims of the Laplace / Normal log-density (distributions.py:150-151)."""
        from . import engine
        return engine.base_log_prob(self, x)

    def sample(self, sample_shape: Optional[Iterable[int]] = None) -> torch.Tensor:
        from . import engine
        return engine.base_sample(self, sample_shape)


class Laplace(DistributionModule):
    base_kind = ops.BASE_LAPLACE

    def __init__(self, loc: torch.Tensor, scale: torch.Tensor, device: str = "cpu"):
        super().__init__()
        self.loc = Parameter(loc)
        self.scale_unconstrained = Parameter(inv_softplus(scale))
        self.to(device)


class Normal(DistributionModule):
    base_kind = ops.BASE_NORMAL

    def __init__(self, loc: torch.Tensor, scale: torch.Tensor, device: str = "cpu"):
        super().__init__()
        self.loc = Parameter(loc)
        self.scale_unconstrained = Parameter(inv_softplus(scale))
        self.to(device)


class Independent(Module):
    """Reinterprets batch dims of a DistributionModule as event dims (distributions.py:709-728).  The fused
    base-density kernel already sums over every non-batch dim, so this is bookkeeping only."""

    def __init__(self, base_distribution: DistributionModule, reinterpreted_batch_ndims: int = 0):
        super().__init__()
        self._base_distribution = base_distribution
        self.reinterpreted_batch_ndims = reinterpreted_batch_ndims

    @property
    def base_dist(self):
        return self._base_distribution

    @property
    def batch_shape(self):
        bs = self._base_distribution.batch_shape
        return torch.Size(bs[:len(bs) - self.reinterpreted_batch_ndims])

    @property
    def event_shape(self):
        bs = self._base_distribution.batch_shape
        return torch.Size(bs[len(bs) - self.reinterpreted_batch_ndims:]) + self._base_distribution.event_shape

    def log_prob(self, x):
        return self._base_distribution.log_prob(x)

    def sample(self, sample_shape=None):
        return self._base_distribution.sample(sample_shape)
