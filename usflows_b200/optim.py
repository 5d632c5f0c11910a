"""SophiaG -- the optimiser `Flow.fit` defaults to in the reference (src/usflows/sophia.py:8-199), restated.

Update rule per parameter (sophia.py:170-199):
    p <- p * (1 - lr * weight_decay)
    m <- beta1 * m + (1 - beta1) * g
    p <- p - lr * sign(m) * min(|m| / (rho * bs * h + 1e-15), 1)
`h` is an EMA of g*g that only `update_hessian()` advances (sophia.py:39-56).  The reference's training loop never
calls it (flows.py:195-203), so h stays 0 and the step degenerates to sign-momentum with step size `lr`
 -- this class reproduces exactly that behaviour, including the optional hessian path.
"""
from __future__ import annotations

import torch
from torch.optim.optimizer import Optimizer


class SophiaG(Optimizer):
    def __init__(self, params, lr=1e-4, betas=(0.965, 0.99), rho=0.04, weight_decay=1e-1, *, maximize: bool = False,
                 capturable: bool = False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if rho < 0.0:
            raise ValueError(f"Invalid rho parameter: {rho}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, rho=rho, weight_decay=weight_decay, maximize=maximize,
                                      capturable=capturable))

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = torch.zeros((), dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["hessian"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def update_hessian(self):
        """h <- beta2 * h + (1 - beta2) * g * g  (sophia.py:39-56)."""
        for group in self.param_groups:
            _, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is not None:
                    self._state(p)["hessian"].mul_(beta2).addcmul_(p.grad, p.grad, value=1 - beta2)

    @torch.no_grad()
    def step(self, closure=None, bs: int = 5120):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, _ = group["betas"]
            lr, rho, wd = group["lr"], group["rho"], group["weight_decay"]
            ps, gs, ms, hs = [], [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("SophiaG does not support sparse gradients")
                st = self._state(p)
                st["step"] += 1
                ps.append(p)
                gs.append(p.grad)
                ms.append(st["exp_avg"])
                hs.append(st["hessian"])
            if not ps:
                continue
            # the same elementwise arithmetic as the per-parameter loop of sophia.py:170-199, issued as multi-tensor
            # ops (one launch per operation for ALL parameters instead of ~8 launches per parameter)
            if group["maximize"]:
                gs = torch._foreach_neg(gs)
            torch._foreach_mul_(ps, 1 - lr * wd)
            torch._foreach_mul_(ms, beta1)
            torch._foreach_add_(ms, gs, alpha=1 - beta1)
            den = torch._foreach_mul(hs, rho * bs)
            torch._foreach_add_(den, 1e-15)
            ratio = torch._foreach_abs(ms)
            torch._foreach_div_(ratio, den)
            torch._foreach_clamp_max_(ratio, 1.0)
            torch._foreach_addcmul_(ps, torch._foreach_sign(ms), ratio, value=-lr)
        return loss
