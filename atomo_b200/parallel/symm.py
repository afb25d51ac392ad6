"""Symmetric NVLink peer-memory heap (Python side).

One :class:`SymmetricHeap` per rank: the same byte size on every rank, every
peer's allocation mapped locally (``ptr(rank)``), optionally an NVLS multicast
alias (``mc_ptr``).  Regions are carved out by name with identical offsets on
all ranks, so ``region_ptr(name, peer)`` is the address of that region inside
any peer — this is what the fused kernels receive instead of MPI tags
(SURVEY.md 5.8).

Creation is collective over the bootstrap process group (gloo or NCCL — used
only for rendezvous: agreeing on the mode and barriers between the multicast
phases).  Mode ``vmm`` = CUDA VMM + POSIX-fd exchange over unix sockets
(``csrc/symm_heap.cpp``) and supports multicast; mode ``ipc`` = legacy cudaIpc
fallback (no multicast); ``world == 1`` is a plain local allocation.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops._ext import load as _load_ext

_HEAP_SEQ = 0


def _all_ok(flag: bool, group=None) -> bool:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flag
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


class SymmetricHeap:
    def __init__(self, nbytes: int, rank: int = 0, world: int = 1, device: Optional[int] = None, group=None,
                 multicast: bool = True, mode: str = "auto", timeout_s: float = 60.0):
        global _HEAP_SEQ
        self.C = _load_ext()
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.cuda.current_device() if device is None else device
        self._regions: Dict[str, Tuple[int, int]] = {}
        self._cursor = 0
        self.h = 0
        _HEAP_SEQ += 1
        job = "%s.%s.%d" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "x"), _HEAP_SEQ)
        want_vmm = mode in ("auto", "vmm") and (world == 1 or self.C.heap_posix_fd_supported(self.device))
        ok = False
        if want_vmm:
            self.h = self.C.heap_create_vmm(rank, world, self.device, int(nbytes), job, bool(multicast), timeout_s)
            ok = _all_ok(self.h != 0, group)
            if not ok:
                err = self.C.heap_last_error()
                if self.h:
                    self.C.heap_destroy(self.h)
                    self.h = 0
                if mode == "vmm":
                    raise RuntimeError("symmetric heap (vmm) failed: %s" % err)
        if not ok:
            self.h, handle = self.C.heap_create_ipc(rank, world, self.device, int(nbytes))
            if not _all_ok(self.h != 0, group):
                raise RuntimeError("symmetric heap (ipc) failed: %s" % self.C.heap_last_error())
            if world > 1:
                handles = [None] * world
                dist.all_gather_object(handles, bytes(handle), group=group)
                if not _all_ok(self.C.heap_open_ipc(self.h, b"".join(handles)), group):
                    raise RuntimeError("cudaIpcOpenMemHandle failed: %s" % self.C.heap_last_error())
        self.mode = self.C.heap_mode(self.h)
        self.nbytes = int(self.C.heap_bytes(self.h))
        # ---- NVLS multicast (optional) ---------------------------------------
        self.mc_error = ""
        if multicast and world > 1 and self.mode == "vmm":
            a = self.C.heap_mc_phase_a(self.h, timeout_s)
            if not a:
                self.mc_error = self.C.heap_last_error()
            if _all_ok(bool(a), group):
                dist.barrier(group=group)
                b = self.C.heap_mc_phase_b(self.h)
                if not b:
                    self.mc_error = self.C.heap_last_error()
                if not _all_ok(bool(b), group):
                    self._mc = 0
                else:
                    self._mc = int(self.C.heap_mc_ptr(self.h))
            else:
                self._mc = 0
        else:
            self._mc = 0
        if world > 1:
            dist.barrier(group=group)

    # ------------------------------------------------------------------
    @property
    def has_multicast(self) -> bool:
        return self._mc != 0

    def ptr(self, rank: Optional[int] = None) -> int:
        return int(self.C.heap_ptr(self.h, self.rank if rank is None else rank))

    def mc_ptr(self) -> int:
        return self._mc

    def alloc(self, name: str, nbytes: int, align: int = 256) -> int:
        """Reserve a named region (same call order on every rank -> same offsets)."""
        off = (self._cursor + align - 1) // align * align
        if off + nbytes > self.nbytes:
            raise MemoryError("symmetric heap exhausted: need %d more bytes" % (off + nbytes - self.nbytes))
        self._regions[name] = (off, nbytes)
        self._cursor = off + nbytes
        return off

    def region_ptr(self, name: str, rank: Optional[int] = None) -> int:
        return self.ptr(rank) + self._regions[name][0]

    def region_mc_ptr(self, name: str) -> int:
        return self._mc + self._regions[name][0] if self._mc else 0

    def tensor(self, name: str, dtype: torch.dtype = torch.float32, rank: Optional[int] = None) -> torch.Tensor:
        """A torch view of a region (local by default; a peer's copy when ``rank`` is given)."""
        off, nbytes = self._regions[name]
        names = {torch.float32: "float32", torch.int32: "int32", torch.int64: "int64", torch.uint8: "uint8",
                 torch.bfloat16: "bfloat16"}
        esize = torch.empty((), dtype=dtype).element_size()
        return self.C.tensor_from_ptr(self.ptr(rank) + off, nbytes // esize, names[dtype], self.device)

    def close(self):
        if self.h:
            torch.cuda.synchronize()
            self.C.heap_destroy(self.h)
            self.h = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
