"""Phase timers.  The reference uses host ``time.time()`` deltas only
(SURVEY.md 5.1), which is wrong for GPU work without a sync.  ``PhaseTimer``
uses CUDA events on the current stream when timing CUDA work and falls back to
``perf_counter`` on CPU; ``max_over_ranks`` implements the "max over ranks"
rule for every multi-GPU number."""
from __future__ import annotations

import time
from contextlib import contextmanager
from typing import Dict

import torch


class PhaseTimer:
    def __init__(self, cuda: bool = False):
        self.cuda = cuda and torch.cuda.is_available()
        self._events: Dict[str, list] = {}
        self._host: Dict[str, float] = {}

    @contextmanager
    def phase(self, name: str):
        if self.cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            try:
                yield
            finally:
                b.record()
                self._events.setdefault(name, []).append((a, b))
        else:
            t0 = time.perf_counter()
            try:
                yield
            finally:
                self._host[name] = self._host.get(name, 0.0) + time.perf_counter() - t0

    def seconds(self) -> Dict[str, float]:
        out = dict(self._host)
        if self.cuda:
            torch.cuda.synchronize()
            for name, pairs in self._events.items():
                out[name] = out.get(name, 0.0) + sum(a.elapsed_time(b) for a, b in pairs) / 1e3
        return out

    def reset(self):
        self._events.clear()
        self._host.clear()


def max_over_ranks(value: float, device=None) -> float:
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
