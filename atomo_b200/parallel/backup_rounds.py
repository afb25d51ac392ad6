"""Backup-worker rounds for the gloo role path: the PS never waits for a straggler and survives a lost worker.

The reference parsed ``--num-aggregate`` and never used it (``/root/reference/src/distributed_nn.py:67``); its PS
blocks forever in ``waitany`` when a worker dies (``sync_replicas_master_nn.py:198-214``).  Protocol
implemented here (PS = rank 0, ``need`` = ``--num-aggregate`` < workers):

* A step is announced point-to-point (step word tag 10 + parameter snapshot tag 11) only to workers that **owe
  nothing**: every gradient message they were asked for has arrived.  A worker that becomes free while the PS is still
  collecting the step (its late message surfaced) is announced the CURRENT step at once, so nobody idles.
* The PS proceeds after ``need`` gradients stamped with the current step.  Late messages carry an old step and are
  dropped; the straggler rejoins with current parameters instead of replaying what it missed.
* All receives run in one receiver thread (any-source, no timeout: a gloo receive that times out tears down every
  connection of the process), and one sentinel receive per worker (tag 99) completes with an error the moment that
  worker's connection closes.  Both feed one queue, so the PS wakes up for "a message arrived" and for "worker w is
  gone" alike; a lost worker is removed from the bookkeeping and ``need`` shrinks to the number of survivors.
* Shutdown: after STOP every worker answers with a bye on both tags, which ends the receiver and the sentinels
  without leaving a pending receive behind.
"""
from __future__ import annotations

import queue
import threading
from typing import Dict, List, Optional, Set

import torch
import torch.distributed as dist

STOP_STEP = -1


class _PayloadLost(RuntimeError):
    def __init__(self, sender, why):
        super().__init__(str(why))
        self.sender = sender


class BackupRounds:
    """PS-side state machine; owned by :class:`~atomo_b200.parallel.transport.TorchDistTransport`."""

    def __init__(self, transport, need: int):
        self.t = transport
        self.need = int(need)
        self.workers = list(range(1, transport.world_size))
        self.sent: Dict[int, int] = {w: 0 for w in self.workers}
        self.recvd: Dict[int, int] = {w: 0 for w in self.workers}
        self.dead: Set[int] = set()
        self.byes: Set[int] = set()
        self.inflight: Dict[int, list] = {w: [] for w in self.workers}   # send requests + tensors kept alive
        self.round = None                     # (step, step tensor, parameter snapshot)
        self.asked: Set[int] = set()          # workers announced the current step
        self.q: "queue.Queue" = queue.Queue()
        self._rx_expect: Optional[Set[int]] = None
        self._threads: List[threading.Thread] = []
        self._spawn(self._rx_loop)
        for w in self.workers:
            self._spawn(self._watch, w)

    # ---- threads ------------------------------------------------------------------------------------------
    def _spawn(self, fn, *a):
        th = threading.Thread(target=fn, args=a, daemon=True)
        th.start()
        self._threads.append(th)

    def _rx_loop(self):
        while True:
            try:
                sender, msg_step, buf = self.t._recv_one(None)
            except _PayloadLost as e:          # the sender died between its header and its payload
                self.q.put(("lost", e.sender, e))
                continue
            except Exception as e:  # every connection is gone (or the process group is being torn down)
                self.q.put(("error", None, e))
                return
            if msg_step == STOP_STEP:
                self.byes.add(sender)
                self.q.put(("bye", sender, None))
                if self._rx_expect is not None and (self._rx_expect - self.dead) <= self.byes:
                    return
                continue
            self.q.put(("msg", sender, (msg_step, buf)))

    def _watch(self, w: int):
        try:
            dist.recv(torch.zeros(1), src=w, group=self.t.group, tag=99)   # completes normally with the bye
        except Exception as e:
            self.q.put(("lost", w, e))

    # ---- bookkeeping --------------------------------------------------------------------------------------
    def live(self) -> List[int]:
        return [w for w in self.workers if w not in self.dead]

    def free(self) -> List[int]:
        return [w for w in self.live() if self.recvd[w] == self.sent[w]]

    def _lost(self, w: int, why) -> None:
        if w not in self.dead:
            self.dead.add(w)
            self.inflight[w] = []
            print("Master: worker {} is gone ({}); continuing with {} workers".format(
                w, str(why).strip().splitlines()[-1][-80:] if why else "connection closed", len(self.live())),
                flush=True)
        if not self.live():
            raise RuntimeError("every worker is gone")

    def _pump(self):
        """Block for the next event.  Returns (sender, msg_step, buf) for a gradient message, else None."""
        kind, w, payload = self.q.get()
        if kind == "lost":
            self._lost(w, payload)
            return None
        if kind == "error":
            raise RuntimeError("backup rounds: the receiver thread failed: %s" % (payload,))
        if kind == "bye":
            return None
        self.recvd[w] += 1
        return (w,) + payload

    def _announce(self, w: int) -> None:
        step, t, snap = self.round
        for r in self.inflight[w]:            # w owes nothing, so it received everything sent before: returns at once
            try:
                r[0].wait()
            except Exception as e:
                return self._lost(w, e)
        try:
            reqs = [(dist.isend(t, dst=w, group=self.t.group, tag=10), t),
                    (dist.isend(snap, dst=w, group=self.t.group, tag=11), snap)]
        except Exception as e:
            return self._lost(w, e)
        self.inflight[w] = reqs
        self.sent[w] += 1
        self.asked.add(w)

    # ---- API used by the PS -------------------------------------------------------------------------------
    def send_round(self, step: int, flat: torch.Tensor) -> List[int]:
        while len(self.free()) < min(self.need, len(self.live())):
            if self._pump() is not None:      # a late message: dropped, its sender is free again
                self.t._stale_dropped += 1
        # the snapshot outlives this call: the PS updates `flat` while a slow receiver may still be reading
        self.round = (step, torch.tensor([step], dtype=torch.int64, device=self.t.device), flat.detach().clone())
        self.asked = set()
        for w in self.free():
            self._announce(w)
        return sorted(self.asked)

    def gather(self, step: int, unpack) -> Dict[int, list]:
        got: Dict[int, list] = {}
        while True:
            outstanding = [w for w in self.asked if w not in got and w not in self.dead]
            required = min(self.need, len(got) + len(outstanding))
            if required == 0:
                required = 1                  # everyone asked is gone: a straggler will join the step below
            if len(got) >= required:
                return got
            m = self._pump()
            if m is None:
                continue
            sender, msg_step, buf = m
            if msg_step == step:
                codes = unpack(buf)["codes"]
                if codes is not None:         # None = abandoned after a kill signal (always stamped with an
                    got[sender] = codes       # older step: the signal is sent after the step's gather returned)
            else:
                self.t._stale_dropped += 1
                if sender not in self.dead:
                    self._announce(sender)    # late joiner: hand it the current step right away

    def finish(self, send_stop) -> int:
        """Collect what is still owed, send STOP, wait for the byes.  Returns the number of dropped messages."""
        dropped = 0
        while any(self.recvd[w] < self.sent[w] for w in self.live()):
            if self._pump() is not None:
                dropped += 1
        self.t._stale_dropped += dropped
        self._rx_expect = set(self.live())
        for w in self.live():
            try:
                send_stop(w)
            except Exception as e:
                self._lost(w, e)
        while not (set(self.live()) <= self.byes):
            self._pump()
        for th in self._threads:
            th.join(timeout=5.0)
        return dropped

    @property
    def clean(self) -> bool:
        """False when a receive is still pending (a worker died): the caller must not tear the group down politely."""
        return not self.dead and not any(th.is_alive() for th in self._threads)
