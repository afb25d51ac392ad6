"""Dense ("sgd") coder — the module the reference forgot to ship.

``/root/reference/src/sync_replicas_master_nn.py:138`` and
``distributed_worker.py:131`` instantiate
``codings.lossless_compress.LosslessCompress()`` for the default ``--code sgd``
path, but the file is absent from the tree.  From its name, the
``_FAKE_SGD`` comment and ``src/utils.py`` it is a pass-through coder that
byte-compresses the raw gradient.  Here: ``encode`` ships the fp32 tensor
as-is (``compress=False``, the GPU/NVLink default — byte-level LZ has no place
on a 900 GB/s link) or, for the CPU/gloo path, as a ``utils.compress``-ed byte
string; ``decode`` restores it exactly.
"""
from __future__ import annotations

import numpy as np
import torch

from .coding import Coding, register
from ..utils.compress import compress as _compress, decompress as _decompress


@register("sgd")
@register("dense")
@register("lossless")
class LosslessCompress(Coding):
    def __init__(self, compress: bool = False, level: int = 1, *args, **kwargs):
        super().__init__()
        self.compress = bool(compress)
        self.level = int(level)

    def encode(self, grad: torch.Tensor, **kwargs) -> dict:
        if not self.compress:
            return {"grad": grad.detach().to(torch.float32), "encode": False, "shape": list(grad.shape)}
        raw = grad.detach().to(torch.float32).cpu().contiguous().numpy().tobytes()
        return {"blob": _compress(raw, level=self.level), "encode": False, "shape": list(grad.shape),
                "compressed": True}

    def decode(self, code: dict, cuda: bool = False, **kwargs) -> torch.Tensor:
        if code.get("compressed", False):
            raw = _decompress(code["blob"])
            g = torch.from_numpy(np.frombuffer(raw, dtype=np.float32).copy()).reshape(code["shape"])
        else:
            g = torch.as_tensor(code["grad"], dtype=torch.float32).reshape(code["shape"])
        return g.cuda() if cuda else g
