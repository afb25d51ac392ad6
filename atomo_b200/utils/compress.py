"""Byte-level lossless compression helpers.

Parity: ``/root/reference/src/utils.py:3-16`` (``compress(msg, level, name)`` /
``decompress(code)`` over python-blosc).  blosc is not available offline, so
the codec is zlib (stdlib); the ``name`` argument is accepted and the same
codecs the reference forbids (``lz4``/``snappy``) are rejected.
"""
import zlib


def compress(msg: bytes, level: int = 0, name: str = "blosclz") -> bytes:
    if name in ("lz4", "snappy"):
        raise ValueError("Do not specify lz4 or snappy. I ran into hard to debug issues")
    return zlib.compress(bytes(msg), level)


def decompress(code: bytes) -> bytes:
    return zlib.decompress(bytes(code))
