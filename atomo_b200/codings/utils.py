"""Atomic-norm indicators used to choose spectral vs entry-wise atoms.

Parity: ``/root/reference/src/codings/utils.py:3-8`` — the ATOMO paper compares
``||G||_* * sqrt(m+n)`` against ``||vec(G)||_1``: whichever is smaller tells
which atomic decomposition yields the lower-variance sparsification.
"""
import math

import torch


def nuclear_indicator(grad: torch.Tensor, s: torch.Tensor) -> float:
    m, n = grad.shape
    return float(s.sum()) * math.sqrt(m + n)


def l1_indicator(grad: torch.Tensor) -> float:
    return float(grad.reshape(-1).abs().sum())


def prefer_spectral(grad2d: torch.Tensor) -> bool:
    """True when spectral atoms are predicted to sparsify with lower variance."""
    s = torch.linalg.svdvals(grad2d.to(torch.float32))
    return nuclear_indicator(grad2d, s) <= l1_indicator(grad2d)
