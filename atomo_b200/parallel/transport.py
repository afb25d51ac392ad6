"""First-class communication layer (the reference has none: MPI calls are
inlined at ~70 sites).

:class:`Transport` is the protocol the PS / worker roles talk to:

====================  =====================================================
reference call site   Transport method
====================  =====================================================
C1 ``isend(step)``    ``send_step(step)``          (master:246-252)
C2 ``irecv(tag 10)``  ``recv_step()``              (worker:267-273)
C3/C4 ``Bcast``       ``bcast_params(flat)``       (master:270-279, worker:275-287)
C7 ``isend(tag 88+i)````push(codes, step)``        (worker:330-335)
C5/C6 ``irecv``/``waitany`` ``gather(step, need)`` (master:281-290, 198-214)
C10 tag 77            ``send_kill(w)`` / ``kill_requested()``  (lenet.py:173-180)
(none)                ``send_round(step, flat, need)`` / ``fetch_params(flat)`` / ``finish()``: backup-worker rounds
====================  =====================================================

Backup workers (``--num-aggregate N`` < workers, gloo): see :mod:`atomo_b200.parallel.backup_rounds` — the PS announces
a step point-to-point only to workers that owe it nothing, proceeds after ``N`` fresh gradients, never waits for a
straggler, and keeps going when a worker's connection closes.

:class:`TorchDistTransport` implements it over ``torch.distributed`` — gloo on
CPU (BASELINE config 1) or NCCL on GPUs (the *baseline* the fused NVLink path
in ``atomo_b200.parallel.symm`` / ``runtime.engine`` is measured against).
Differences from the reference: parameters travel as ONE flat fp32 broadcast
(not P float64 Bcasts), a worker's step message is ONE packed buffer
(``wire.pack``), and messages carry the step so stale gradients are dropped
(the reference's ``generate_tag`` idea, resnet_split.py:25-39, made real).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import wire

STOP_STEP = -1


class Transport:
    rank: int
    world_size: int

    @property
    def num_workers(self) -> int:
        return self.world_size - 1

    def send_step(self, step: int) -> None: raise NotImplementedError
    def recv_step(self) -> int: raise NotImplementedError
    def bcast_params(self, flat: torch.Tensor) -> None: raise NotImplementedError
    def push(self, codes: list, step: int) -> int: raise NotImplementedError
    def gather(self, step: int, need: Optional[int] = None): raise NotImplementedError
    def send_kill(self, worker_rank: int) -> None: raise NotImplementedError
    def enable_backup_rounds(self, need: int) -> bool: return False
    def send_round(self, step: int, flat: torch.Tensor, need: Optional[int] = None) -> List[int]:
        self.send_step(step)
        self.bcast_params(flat)
        return list(range(1, self.world_size))
    def fetch_params(self, flat: torch.Tensor) -> None: self.bcast_params(flat)
    def asked_workers(self) -> List[int]: return list(range(1, self.world_size))
    def lost_workers(self) -> List[int]: return []
    def finish(self) -> None:
        self.drain()
        self.send_step(STOP_STEP)
    backup_rounds = False
    clean_shutdown = True
    def drain(self) -> int: return 0
    def kill_requested(self, step=None) -> bool: return False
    def barrier(self) -> None: raise NotImplementedError


class TorchDistTransport(Transport):
    """PS = rank 0, workers = ranks 1..W-1, over an initialized process group."""

    def __init__(self, device: Optional[torch.device] = None, group=None, timeout_s: float = 300.0):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialized first")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")
        self.device = device
        self.timeout_s = timeout_s
        self._stale_dropped = 0
        self._rounds = 0                      # steps broadcast so far (PS side)
        self._recv_count: Dict[int, int] = {}  # messages received per worker (PS side)
        self._p2p_rounds = False               # backup-worker rounds (parallel/backup_rounds.py)
        self._backup = None
        self._bye_on_stop = False
        self.bytes_sent = 0

    # -- step handshake -------------------------------------------------
    def send_step(self, step: int) -> None:
        if step != STOP_STEP:
            self._rounds += 1
        t = torch.tensor([step], dtype=torch.int64, device=self.device)
        reqs = [dist.isend(t, dst=w, group=self.group, tag=10) for w in range(1, self.world_size)]
        for r in reqs:
            r.wait()

    def recv_step(self) -> int:
        t = torch.zeros(1, dtype=torch.int64, device=self.device)
        dist.recv(t, src=0, group=self.group, tag=10)
        step = int(t.item())
        if step == STOP_STEP and self._p2p_rounds:
            # backup rounds: answer STOP with a bye on the receiver thread's channel (the sentinel's follows)
            dist.send(torch.tensor([0, STOP_STEP], dtype=torch.int64, device=self.device), dst=0, group=self.group, tag=87)
        if step == STOP_STEP and self._bye_on_stop:
            dist.send(torch.zeros(1), dst=0, group=self.group, tag=99)     # completes the PS's sentinel receive
        return step

    # -- parameters -------------------------------------------------------
    def bcast_params(self, flat: torch.Tensor) -> None:
        dist.broadcast(flat, src=0, group=self.group)

    # -- backup-worker rounds (parallel/backup_rounds.py) ---------------------
    def enable_backup_rounds(self, need: int) -> bool:
        """Both roles call this with the same ``need`` before training.  True when steps will be announced
        point-to-point (gloo and 0 < need < workers); NCCL has no any-source receive and keeps the collective
        round, which waits for every worker."""
        self._p2p_rounds = self.backend == "gloo" and 0 < int(need) < self.num_workers
        self._bye_on_stop = self.backend == "gloo"
        if self._p2p_rounds and self.rank == 0:
            from .backup_rounds import BackupRounds
            self._backup = BackupRounds(self, need)
        elif self.backend == "gloo" and self.rank == 0:
            self._start_fail_fast_watch()
        return self._p2p_rounds

    def _start_fail_fast_watch(self) -> None:
        """All-workers mode on gloo: the PS cannot continue without a worker (collective broadcast, gather of W
        gradients), and a blocked any-source receive would only notice after the process-group timeout (30 min; the
        reference: never, master:198-214).  One sentinel receive per worker completes with an error the moment that
        worker's connection closes: report it and stop the job at once."""
        import os
        import sys
        import threading

        def watch(w):
            try:
                dist.recv(torch.zeros(1), src=w, group=self.group, tag=99)     # completes normally with the bye
            except Exception as e:
                print("Master: worker {} is gone ({}): stopping the job.  With --num-aggregate N < workers the PS "
                      "keeps training with the survivors.".format(w, str(e).strip().splitlines()[-1][-80:]), flush=True)
                sys.stderr.flush()
                os._exit(3)

        for w in range(1, self.world_size):
            threading.Thread(target=watch, args=(w,), daemon=True).start()

    @property
    def backup_rounds(self) -> bool:
        return self._p2p_rounds

    def send_round(self, step: int, flat: torch.Tensor, need: Optional[int] = None) -> List[int]:
        """PS: announce ``step`` and its parameters.  Returns the worker ranks that were asked for a gradient."""
        if not self._p2p_rounds:
            return super().send_round(step, flat, need)
        return self._backup.send_round(step, flat)

    def fetch_params(self, flat: torch.Tensor) -> None:
        """Worker: receive the parameters of the step just announced."""
        if self._p2p_rounds:
            dist.recv(flat, src=0, group=self.group, tag=11)
        else:
            self.bcast_params(flat)

    def asked_workers(self) -> List[int]:
        """Workers that were announced the current step (late joiners included)."""
        return sorted(self._backup.asked) if self._p2p_rounds else list(range(1, self.world_size))

    def lost_workers(self) -> List[int]:
        return sorted(self._backup.dead) if self._p2p_rounds and self.rank == 0 else []

    def finish(self) -> None:
        """PS: end of training — collect what stragglers still owe, then STOP every worker."""
        if not self._p2p_rounds:
            self.drain()
            return self.send_step(STOP_STEP)
        stop = torch.tensor([STOP_STEP], dtype=torch.int64, device=self.device)
        self._backup.finish(lambda w: dist.send(stop, dst=w, group=self.group, tag=10))

    @property
    def clean_shutdown(self) -> bool:
        """False when a worker was lost: a receive on its connection is still pending, skip the polite teardown."""
        return self._backup.clean if self._p2p_rounds and self.rank == 0 else True

    # -- gradients: worker side ------------------------------------------
    def push(self, codes: list, step: int) -> int:
        buf = wire.pack({"step": step, "rank": self.rank, "codes": codes}, device=self.device)
        n = torch.tensor([buf.numel(), step], dtype=torch.int64, device=self.device)
        dist.send(n, dst=0, group=self.group, tag=87)
        dist.send(buf, dst=0, group=self.group, tag=88)
        self.bytes_sent += buf.numel()
        return buf.numel()

    # -- gradients: PS side ------------------------------------------------
    def _recv_one(self, src):
        """Receive one (header, payload) message; ``src=None`` = any source
        (the reference's ``waitany``, master:198-214 — gloo only)."""
        hdr = torch.zeros(2, dtype=torch.int64, device=self.device)
        sender = dist.recv(hdr, src=src, group=self.group, tag=87)
        sender = src if src is not None else sender
        nbytes, msg_step = int(hdr[0].item()), int(hdr[1].item())
        if nbytes == 0:                       # header-only message (the bye of a backup round)
            return sender, msg_step, None
        buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        try:
            dist.recv(buf, src=sender, group=self.group, tag=88)
        except Exception as e:
            if self._p2p_rounds:
                from .backup_rounds import _PayloadLost
                raise _PayloadLost(sender, e)
            raise
        return sender, msg_step, buf

    def gather(self, step: int, need: Optional[int] = None):
        """Collect this step's messages.  Returns ``{worker_rank: codes}``.

        ``need`` < num_workers gives backup-worker semantics (the reference's
        ``--num-aggregate``, parsed at launcher:67 but never used): return as
        soon as ``need`` workers have delivered (arrival order on gloo via
        any-source receive; rank order on NCCL, which has no any-source).
        Stragglers' messages carry an old step and are dropped when they
        surface in a later call.  A dead worker surfaces as the process-group
        timeout instead of the reference's infinite ``waitany`` block.
        """
        if self._p2p_rounds:
            return self._backup.gather(step, wire.unpack)
        need = self.num_workers if need is None else min(need, self.num_workers)
        got: Dict[int, list] = {}
        any_source = self.backend == "gloo"
        order = list(range(1, self.world_size))
        if not any_source:
            need = self.num_workers  # NCCL has no any-source receive: wait for everyone, rank order
        cursor = 0
        while len(got) < need:
            if any_source:
                sender, msg_step, buf = self._recv_one(None)
            else:
                sender, msg_step, buf = self._recv_one(order[cursor])
                cursor += 1
            self._recv_count[sender] = self._recv_count.get(sender, 0) + 1
            if msg_step == step:
                codes = wire.unpack(buf)["codes"]
                if codes is None:      # the worker abandoned this step after a kill signal
                    continue
                got[sender] = codes
            else:  # stale gradient from a straggler: drop
                self._stale_dropped += 1
        return got

    def drain(self) -> int:
        """PS side, before STOP: receive (and drop) the messages stragglers still owe
        us, so no worker is left blocked in a send when the PS exits."""
        dropped = 0
        for w in range(1, self.world_size):
            while self._recv_count.get(w, 0) < self._rounds:
                self._recv_one(w)
                self._recv_count[w] = self._recv_count.get(w, 0) + 1
                dropped += 1
        self._stale_dropped += dropped
        return dropped

    # -- straggler kill signal (the reference's tag 77) ------------------------
    # The reference polls ``Iprobe(0, 77)`` between layer backwards (lenet.py:173-180).  gloo has no
    # probe and a blocking listener thread fights the compute threads, so the signal travels through
    # the rendezvous key-value store instead: a step-stamped key per worker, checked without blocking.
    def _store(self):
        if getattr(self, "_kv", None) is None:
            try:
                from torch.distributed.distributed_c10d import _get_default_store
                self._kv = _get_default_store()
            except Exception:
                self._kv = False
        return None if self._kv is False else self._kv

    def send_kill(self, worker_rank: int, step: int = 0) -> None:
        kv = self._store()
        if kv is not None:
            kv.set("atomo_b200/kill/%d" % worker_rank, str(step))

    def enable_kill_listener(self) -> None:
        self._kill_enabled = self._store() is not None

    def kill_requested(self, step: Optional[int] = None) -> bool:
        if not getattr(self, "_kill_enabled", False):
            return False
        kv, key = self._store(), "atomo_b200/kill/%d" % self.rank
        try:
            if not kv.check([key]):
                return False
            stamped = int(kv.get(key))
        except Exception:
            return False
        # signals are step-stamped: a late one for step t cannot abort step t+1
        return step is None or stamped == step

    def barrier(self) -> None:
        dist.barrier(group=self.group)
