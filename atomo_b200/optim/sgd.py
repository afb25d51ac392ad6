"""Parameter-server momentum SGD with an *external gradient list* API.

Parity: ``/root/reference/src/optim/sgd.py:57-90`` — ``step(grads, closure,
cuda)`` takes one gradient per parameter (in ``parameters()`` order) because
the PS never runs backward: the gradients are the decoded, averaged worker
messages.  Semantics are PyTorch's: L2 weight decay, ``buf = mu*buf +
(1-dampening)*g`` (first step ``buf = g``), optional Nesterov, ``p -= lr*d``.

Differences: gradients are torch tensors (numpy arrays are accepted and
converted), ``set_lr`` actually reaches the param groups (the reference's LR
decay never did), and on CUDA the same update is available fused
into the PS decode kernel (``atomo_b200.ops.ps_update``) — this class is then
only the state container/oracle.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
from torch.optim.optimizer import Optimizer, required


def _as_tensor(g, like: torch.Tensor) -> torch.Tensor:
    if not isinstance(g, torch.Tensor):
        g = torch.as_tensor(g)
    return g.to(device=like.device, dtype=like.dtype).reshape(like.shape)


class SGD(Optimizer):
    def __init__(self, params, lr=required, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        if lr is not required and lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if momentum < 0.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening,
                        weight_decay=weight_decay, nesterov=nesterov)
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
            wd, mu = group["weight_decay"], group["momentum"]
            damp, nesterov, lr = group["dampening"], group["nesterov"], group["lr"]
            for p in group["params"]:
                if grads is None:
                    if p.grad is None:
                        offset += 1
                        continue
                    d_p = p.grad
                else:
                    d_p = _as_tensor(grads[offset], p)
                offset += 1
                if wd != 0:
                    d_p = d_p.add(p, alpha=wd)
                if mu != 0:
                    state = self.state[p]
                    if "momentum_buffer" not in state:
                        buf = state["momentum_buffer"] = torch.clone(d_p).detach()
                    else:
                        buf = state["momentum_buffer"]
                        buf.mul_(mu).add_(d_p, alpha=1 - damp)
                    d_p = d_p.add(buf, alpha=mu) if nesterov else buf
                p.add_(d_p, alpha=-lr)
        return loss
