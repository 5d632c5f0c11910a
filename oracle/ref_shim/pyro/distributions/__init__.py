import torch
from torch.distributions import *  # noqa: F401,F403
from torch.distributions import constraints, Distribution, Independent  # noqa: F401
from . import transforms  # noqa: F401


class TransformModule(torch.distributions.Transform, torch.nn.Module):
    """Same definition as pyro.distributions.TransformModule (a Transform that is also an nn.Module)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def __hash__(self):
        return super(torch.nn.Module, self).__hash__()


class TransformedDistribution(torch.distributions.TransformedDistribution):
    def clear_cache(self):  # pyro adds this; reference Flow.fit calls it (flows.py:207)
        pass
