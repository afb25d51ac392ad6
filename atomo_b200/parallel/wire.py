"""Fixed binary wire format for coded gradients.

The reference pickles a Python dict per layer (``distributed_worker.py:324``,
``sync_replicas_master_nn.py:323``).  Here a step's codes (one per parameter)
are packed into ONE contiguous ``uint8`` tensor: a small JSON header
(metadata, tensor dtypes/shapes/offsets) followed by 16-byte-aligned raw tensor
payloads — one message per worker per step instead of P pickles, and directly
sendable with ``torch.distributed`` (gloo or NCCL) without Python object
serialization of tensor data.
"""
from __future__ import annotations

import json
import struct
from typing import Any, Dict, List

import numpy as np
import torch

_MAGIC = 0xA70B200
_ALIGN = 16


def _flatten(obj: Any, tensors: List[torch.Tensor]):
    if isinstance(obj, torch.Tensor):
        tensors.append(obj)
        return {"__t__": len(tensors) - 1}
    if isinstance(obj, (bytes, bytearray)):
        tensors.append(torch.frombuffer(bytearray(obj), dtype=torch.uint8))
        return {"__b__": len(tensors) - 1}
    if isinstance(obj, dict):
        return {"__d__": {k: _flatten(v, tensors) for k, v in obj.items()}}
    if isinstance(obj, (list, tuple)):
        return {"__l__": [_flatten(v, tensors) for v in obj]}
    if isinstance(obj, torch.Size):
        return {"__l__": [int(v) for v in obj]}
    if isinstance(obj, (int, float, str, bool)) or obj is None:
        return obj
    if isinstance(obj, np.generic):
        return obj.item()
    raise TypeError("cannot serialize %r" % type(obj))


def _unflatten(obj: Any, tensors: List[torch.Tensor]):
    if isinstance(obj, dict):
        if "__t__" in obj:
            return tensors[obj["__t__"]]
        if "__b__" in obj:
            return bytes(tensors[obj["__b__"]].cpu().numpy().tobytes())
        if "__d__" in obj:
            return {k: _unflatten(v, tensors) for k, v in obj["__d__"].items()}
        if "__l__" in obj:
            return [_unflatten(v, tensors) for v in obj["__l__"]]
    return obj


def pack(obj: Any, device=None) -> torch.Tensor:
    """Serialize a nested structure of tensors/metadata into a uint8 tensor."""
    tensors: List[torch.Tensor] = []
    tree = _flatten(obj, tensors)
    metas, offset = [], 0
    for t in tensors:
        nbytes = t.numel() * t.element_size()
        metas.append({"dtype": str(t.dtype).replace("torch.", ""), "shape": list(t.shape),
                      "offset": offset, "nbytes": nbytes})
        offset += (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    header = json.dumps({"tree": tree, "tensors": metas}).encode()
    hpad = (len(header) + _ALIGN - 1) // _ALIGN * _ALIGN
    dev = device if device is not None else (tensors[0].device if tensors else torch.device("cpu"))
    out = torch.zeros(16 + hpad + offset, dtype=torch.uint8, device=dev)
    prefix = struct.pack("<IIQ", _MAGIC, len(header), offset)
    out[:16] = torch.frombuffer(bytearray(prefix), dtype=torch.uint8).to(dev)
    out[16:16 + len(header)] = torch.frombuffer(bytearray(header), dtype=torch.uint8).to(dev)
    base = 16 + hpad
    for t, m in zip(tensors, metas):
        if m["nbytes"]:
            src = t.detach().contiguous().reshape(-1).view(torch.uint8).to(dev)
            out[base + m["offset"]: base + m["offset"] + m["nbytes"]] = src
    return out


def unpack(buf: torch.Tensor) -> Any:
    head = bytes(buf[:16].cpu().numpy().tobytes())
    magic, hlen, _ = struct.unpack("<IIQ", head)
    if magic != _MAGIC:
        raise ValueError("bad wire magic")
    header = json.loads(bytes(buf[16:16 + hlen].cpu().numpy().tobytes()).decode())
    hpad = (hlen + _ALIGN - 1) // _ALIGN * _ALIGN
    base = 16 + hpad
    tensors = []
    for m in header["tensors"]:
        dtype = getattr(torch, m["dtype"])
        raw = buf[base + m["offset"]: base + m["offset"] + m["nbytes"]]
        if m["nbytes"] == 0:
            tensors.append(torch.zeros(m["shape"], dtype=dtype, device=buf.device))
        else:
            tensors.append(raw.clone().view(dtype).reshape(m["shape"]))
    return _unflatten(header["tree"], tensors)
