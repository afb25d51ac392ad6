"""Entry-wise ATOMO: standard-basis atoms.

The reference ships only the indicator for this decomposition
(``/root/reference/src/codings/utils.py:7-8``) and describes the scheme in
``README.md:5-7``; BASELINE.json configs 1 and 5 require it as a coder.  Atoms
are the tensor's entries: ``p_i = min(1, s*|g_i|/||g||_1)`` (optionally
water-filled), keep ``g_i/p_i`` w.p. ``p_i``.  The code is a compacted
``(idx int32, val fp32)`` list — the layout ``csrc/entrywise_kernels.cu``
streams into PS peer memory.

``budget`` is either an absolute expected atom count (``budget >= 1``) or a
fraction of ``numel`` (``0 < budget < 1``; e.g. 0.01 / 0.05 / 0.25 as in the
BASELINE bandwidth sweep).
"""
from __future__ import annotations

from typing import Optional

import torch

from .coding import Coding, register
from .sampling import atom_probabilities


@register("entrywise")
class EntryWise(Coding):
    def __init__(self, budget: float = 0.05, prob_rule: str = "reference", scheme: str = "bernoulli",
                 generator: Optional[torch.Generator] = None, *args, **kwargs):
        super().__init__()
        if budget <= 0:
            raise ValueError("budget must be positive")
        self.budget = float(budget)
        self.prob_rule = prob_rule
        self.scheme = scheme
        self.generator = generator

    def atoms_for(self, numel: int) -> float:
        s = self.budget * numel if self.budget < 1.0 else self.budget
        return float(min(max(s, 1.0), numel))

    def probabilities(self, flat: torch.Tensor) -> torch.Tensor:
        s = self.atoms_for(flat.numel())
        mag = flat.abs().to(torch.float32)
        if self.prob_rule == "reference":
            total = mag.sum()
            if float(total) <= 0:
                return torch.zeros_like(mag)
            return (s * mag / total).clamp(max=1.0)
        return atom_probabilities(mag, s, "waterfill").to(torch.float32)

    def encode(self, grad: torch.Tensor, uniforms: Optional[torch.Tensor] = None, **kwargs) -> dict:
        shape = list(grad.shape)
        flat = grad.detach().reshape(-1).to(torch.float32)
        p = self.probabilities(flat)
        if self.scheme == "systematic":
            u = float(uniforms.flatten()[0]) if uniforms is not None else float(
                torch.rand(1, generator=self.generator))
            c = torch.cumsum(p.to(torch.float64), 0)
            hi = torch.floor(c + u)
            lo = torch.floor(torch.cat([c.new_zeros(1), c[:-1]]) + u)
            keep = hi > lo
        else:
            if uniforms is None:
                dice = torch.rand(flat.shape, generator=self.generator, dtype=torch.float32).to(flat.device)
            else:
                dice = uniforms.reshape(-1)[: flat.numel()].to(flat.device, torch.float32)
            keep = dice < p
        idx = torch.nonzero(keep).flatten()
        val = flat[idx] / p[idx]
        return {"idx": idx.to(torch.int32), "val": val, "shape": shape, "numel": flat.numel()}

    def decode(self, code: dict, cuda: bool = False, **kwargs) -> torch.Tensor:
        out = torch.zeros(int(code["numel"]), dtype=torch.float32, device=code["val"].device)
        out.index_add_(0, code["idx"].to(torch.long), code["val"])
        out = out.reshape(code["shape"])
        return out.cuda() if cuda else out
