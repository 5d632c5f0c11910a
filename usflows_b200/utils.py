import torch


def inv_softplus(x: torch.Tensor) -> torch.Tensor:
    """log(exp(x) - 1): parameterisation of the base-distribution scale (reference utils.py:3-9)."""
    return torch.log(torch.exp(x) - 1)
