"""Coder plug-in base class + registry.

Parity: ``/root/reference/src/codings/coding.py:3-11`` (``Coding.encode`` /
``Coding.decode``) and the registry role of ``src/codings/__init__.py:1-6``.
The new framework adds a name->class registry so the launcher's ``--code`` flag
resolves coders without ``if/elif`` chains, and a ``wire_bytes`` hook used for
the ``Msg(MB)`` column of the worker log line (``distributed_worker.py:255-258``).
"""
from __future__ import annotations

from typing import Callable, Dict, Type

import torch

_REGISTRY: Dict[str, Callable[..., "Coding"]] = {}


def register(name: str):
    """Class decorator: make a coder constructible through :func:`build`."""

    def deco(cls):
        _REGISTRY[name.lower()] = cls
        cls.registry_name = name.lower()
        return cls

    return deco


def available():
    return sorted(_REGISTRY)


def build(name: str, **kwargs) -> "Coding":
    """Instantiate a coder by its ``--code`` name."""
    key = name.lower()
    if key not in _REGISTRY:
        raise ValueError(
            "args.code not recognized: %r (available: %s)" % (name, ", ".join(available()))
        )
    return _REGISTRY[key](**kwargs)


class Coding:
    """Abstract gradient coder.

    ``encode(grad) -> dict`` produces a *code* (a dict of tensors + metadata);
    ``decode(code) -> torch.Tensor`` reconstructs a dense gradient estimate with
    the original shape.  Stochastic coders must be unbiased:
    ``E[decode(encode(g))] == g``.
    """

    registry_name = "coding"

    def __init__(self, *args, **kwargs):
        self.codes = []

    def encode(self, grad: torch.Tensor, *args, **kwargs) -> dict:
        raise NotImplementedError()

    def decode(self, code: dict, *args, **kwargs) -> torch.Tensor:
        raise NotImplementedError()

    # ------------------------------------------------------------------
    @staticmethod
    def wire_bytes(code: dict) -> int:
        """Bytes this code occupies on the wire (tensor payloads only)."""
        total = 0
        for v in code.values():
            if isinstance(v, torch.Tensor):
                total += v.numel() * v.element_size()
            elif isinstance(v, (bytes, bytearray)):
                total += len(v)
            elif isinstance(v, dict):
                total += Coding.wire_bytes(v)
            elif isinstance(v, (list, tuple)):
                for item in v:
                    if isinstance(item, dict):
                        total += Coding.wire_bytes(item)
                    elif isinstance(item, torch.Tensor):
                        total += item.numel() * item.element_size()
        return total
