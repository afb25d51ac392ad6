"""QSVD: jointly sparsify (spectral ATOMO) and quantize (QSGD) the factors.

The reference ships this coder only as stale bytecode
(``/root/reference/src/codings/__pycache__/qsvd.cpython-36.pyc``: ``class
QSVD(Coding)`` holding an ``SVD`` and a ``QSGD`` coder; encode = SVD-encode
then quantize the factors, decode = de-quantize then SVD-decode) and names it
as future work in ``README.md:141-142``.  Both stages are unbiased and
independent, so the composition is unbiased.
"""
from __future__ import annotations

import torch

from .coding import Coding, register
from .qsgd import QSGD
from .svd import SVD


@register("qsvd")
class QSVD(Coding):
    def __init__(self, scheme: str = "qsgd", rank: int = 0, quantization_level: int = 4,
                 bucket_size: int = 512, random_sample: bool = True, *args, **kwargs):
        super().__init__()
        if scheme not in ("qsgd", "terngrad"):
            raise ValueError("scheme must be 'qsgd' or 'terngrad'")
        self.svd = SVD(compress=True, rank=rank, random_sample=random_sample, **kwargs)
        self.quant = QSGD(scheme=scheme, bucket_size=bucket_size, quantization_level=quantization_level)

    def encode(self, grad: torch.Tensor, **kwargs) -> dict:
        code = self.svd.encode(grad, **kwargs)
        return {
            "u": self.quant.encode(code["u"]),
            "vT": self.quant.encode(code["vT"]),
            "s": code["s"],
            "orig_size": code["orig_size"],
            "reshaped": code["reshaped"],
            "encode": True,
            "rank": code["rank"],
        }

    def decode(self, code: dict, cuda: bool = False, **kwargs) -> torch.Tensor:
        inner = {
            "u": self.quant.decode(code["u"]),
            "vT": self.quant.decode(code["vT"]),
            "s": code["s"],
            "orig_size": code["orig_size"],
            "reshaped": code["reshaped"],
            "encode": True,
        }
        return self.svd.decode(inner, cuda=cuda)
