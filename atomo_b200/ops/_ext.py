"""Loader for the in-tree native extension ``atomo_b200._C``.

The extension is built by ``python setup.py build_ext --inplace`` (or
``__graft_entry__.build()``).  On a machine with a GPU a missing extension is a
hard error — the CUDA ops never fall back silently to eager PyTorch.
"""
from __future__ import annotations

import importlib

_C = None
_ERR = None


def load(required: bool = True):
    global _C, _ERR
    if _C is not None:
        return _C
    try:
        _C = importlib.import_module("atomo_b200._C")
    except Exception as e:  # pragma: no cover - depends on the build state
        _ERR = e
        if required:
            raise RuntimeError(
                "atomo_b200 native extension is not built/loadable (%s). Run "
                "`python setup.py build_ext --inplace` in the repo root." % (e,)) from e
        return None
    return _C


def available() -> bool:
    return load(required=False) is not None
