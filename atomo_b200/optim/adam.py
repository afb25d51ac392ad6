"""Parameter-server Adam/AMSGrad with the external-gradient API.

Parity: ``/root/reference/src/optim/adam.py:37-94`` (``step(grads, closure)``,
AMSGrad option, bias-corrected step size, L2 weight decay).  Unlike the
reference it also runs on CUDA parameters.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch.optim.optimizer import Optimizer

from .sgd import _as_tensor


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {}".format(betas))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)

    def set_lr(self, lr: float) -> None:
        for group in self.param_groups:
            group["lr"] = lr

    @torch.no_grad()
    def step(self, grads: Sequence = None, closure=None, cuda: bool = False):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if grads is not None:
            grads = list(grads)
        offset = 0
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if grads is None:
                    if p.grad is None:
                        offset += 1
                        continue
                    grad = p.grad
                else:
                    grad = _as_tensor(grads[offset], p)
                offset += 1
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p)
                    state["exp_avg_sq"] = torch.zeros_like(p)
                    if group["amsgrad"]:
                        state["max_exp_avg_sq"] = torch.zeros_like(p)
                state["step"] += 1
                if group["weight_decay"] != 0:
                    grad = grad.add(p, alpha=group["weight_decay"])
                exp_avg, exp_avg_sq = state["exp_avg"], state["exp_avg_sq"]
                exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
                exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                if group["amsgrad"]:
                    torch.maximum(state["max_exp_avg_sq"], exp_avg_sq, out=state["max_exp_avg_sq"])
                    denom = state["max_exp_avg_sq"].sqrt().add_(group["eps"])
                else:
                    denom = exp_avg_sq.sqrt().add_(group["eps"])
                bc1 = 1 - beta1 ** state["step"]
                bc2 = 1 - beta2 ** state["step"]
                step_size = group["lr"] * math.sqrt(bc2) / bc1
                p.addcdiv_(exp_avg, denom, value=-step_size)
        return loss
