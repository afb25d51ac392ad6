"""Worker role.

Parity: ``DistributedWorker`` (``/root/reference/src/distributed_worker.py:98-370``):
``__init__(comm, **kwargs)``, ``build_model(num_classes)``,
``train(train_loader, test_loader)``; per step: fetch step -> fetch weights ->
forward/backward -> encode every parameter's gradient -> push -> log line
(``worker:255-258``, the format ``tiny_tuning_parser`` greps) -> periodic test
evaluation (``worker:344-370``).  BN running statistics stay private to each
worker like the reference (``worker:301-302``).

Fixes: the CPU path encodes the real gradient (the reference's never assigns
it, worker:319-323); the whole step's codes travel as one message.
"""
from __future__ import annotations

import os
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..models import build_model
from ..parallel.transport import Transport, STOP_STEP
from ..utils import checkpoint as ckpt
from ..utils.logging import MetricsWriter, worker_line, test_line
from .flat import FlatLayout, bind_parameters
from .master import STEP_START_, build_coder
from .nn_ops import NN_Trainer, accuracy
from ..codings import Coding


class DistributedWorker(NN_Trainer):
    def __init__(self, comm: Transport, **kwargs):
        self.comm = comm
        self.world_size = comm.world_size
        self.rank = comm.rank
        self.cur_step = 0
        self.next_step = 0
        self.batch_size = kwargs.get("batch_size", 128)
        self.max_epochs = kwargs.get("max_epochs", 100)
        self.momentum = kwargs.get("momentum", 0.5)
        self.lr = kwargs.get("learning_rate", 0.01)
        self.network_config = kwargs["network"]
        self.dataset = kwargs.get("dataset", "")
        self._max_steps = kwargs.get("max_steps", 10000)
        self.comm_type = kwargs.get("comm_method", "Bcast")
        self._compress = kwargs.get("compress", False)
        self._enable_gpu = bool(kwargs.get("enable_gpu", False)) and torch.cuda.is_available()
        self._eval_batch_size = 100
        self._eval_freq = kwargs.get("eval_freq", 50)
        self._train_dir = kwargs.get("train_dir", "output/models/")
        self._svd_rank = kwargs.get("svd_rank", 0)
        self._quantization_level = kwargs.get("quantization_level", 4)
        self._bucket_size = kwargs.get("bucket_size", 512)
        self._code = kwargs.get("code", "sgd")
        self._eval_batches = kwargs.get("eval_batches", None)
        # layer-wise backward with gradient emission + straggler kill (the reference's *Split models,
        # resnet_split.py:458-570): encode each layer as soon as its gradient exists, abandon the step when
        # the PS signals (tag 77) that it already has enough gradients
        self._split_backward = bool(kwargs.get("split_backward", False))
        self.device = torch.device("cuda", torch.cuda.current_device()) if self._enable_gpu else torch.device("cpu")
        self._coder = build_coder(kwargs, worker_side=True)
        self.last_stats = {}
        self._metrics = MetricsWriter(kwargs.get("metrics_file", ""), self.rank, "worker")
        # fault injection for the backup-worker tests: ATOMO_DEBUG_SLOW_WORKER="<rank>:<seconds per step>"
        slow = os.environ.get("ATOMO_DEBUG_SLOW_WORKER", "")
        self._debug_slow_s = float(slow.split(":")[1]) if slow and int(slow.split(":")[0]) == self.rank else 0.0
        # ATOMO_DEBUG_DIE_WORKER="<rank>:<step>[,<rank>:<step>...]": the process exits in the middle of the first step
        # >= <step> it takes part in (a backup-mode straggler may never see step <step> itself)
        self._debug_die_step = 0
        for item in filter(None, os.environ.get("ATOMO_DEBUG_DIE_WORKER", "").split(",")):
            r, st = item.split(":")
            if int(r) == self.rank:
                self._debug_die_step = int(st)

    def _die_armed(self) -> bool:
        """ATOMO_DEBUG_DIE_ONCE=<marker file>: the injected death happens once per marker (restart tests)."""
        marker = os.environ.get("ATOMO_DEBUG_DIE_ONCE", "")
        if not marker:
            return True
        if os.path.exists(marker):
            return False
        open(marker, "w").close()
        return True

    def build_model(self, num_classes: int = 10):
        self.network = build_model(self.network_config, num_classes, self.dataset)
        if self._split_backward:
            from ..models.split import FC_NN_Split, LeNetSplit, ResNetSplit18, ResNetSplit34
            split = {"LeNet": LeNetSplit, "FC": FC_NN_Split, "ResNet18": ResNetSplit18, "ResNet34": ResNetSplit34}
            if self.network_config not in split:
                raise ValueError("split backward is available for %s" % ", ".join(sorted(split)))
            self.network = split[self.network_config](num_classes=num_classes)
            self.comm.enable_kill_listener()
        self.network = self.network.to(self.device)
        self.layout = FlatLayout.from_module(self.network)
        # the receive buffer IS the parameter storage (parity role: ModelBuffer, worker:84-95)
        self.flat_params = torch.zeros(self.layout.total, dtype=torch.float32, device=self.device)
        bind_parameters(self.network, self.flat_params, self.layout)
        self.optimizer = torch.optim.SGD(self.network.parameters(), lr=self.lr, momentum=self.momentum)
        self.criterion = nn.CrossEntropyLoss()
        return self

    # ------------------------------------------------------------------
    def train(self, train_loader, test_loader=None):
        n_data = len(train_loader.dataset)
        print("Worker {}: starting training".format(self.rank))
        iter_start = time.time()
        for num_epoch in range(self.max_epochs):
            for batch_idx, (x, y) in enumerate(train_loader):
                self.next_step = self.async_fetch_step()
                if self.next_step == STOP_STEP:
                    return
                self.update_step()
                if num_epoch == 0 and batch_idx == 0:
                    assert self.cur_step >= STEP_START_
                iter_start = time.time()
                x, y = x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True)

                t0 = time.time()
                self.async_fetch_weights_bcast()
                fetch_weight_duration = time.time() - t0

                self.network.train()
                self.optimizer.zero_grad()
                comp_start = time.time()
                logits = self.network(x)
                loss = self.criterion(logits, y)
                if self._debug_slow_s:
                    time.sleep(self._debug_slow_s)
                if self._debug_die_step and self.cur_step >= self._debug_die_step and self._die_armed():
                    print("Worker {}: fault injection, dying at step {}".format(self.rank, self.cur_step), flush=True)
                    os._exit(0)
                if self._split_backward:
                    emitted = {}
                    killed = self.network.backward_signal_kill(
                        loss, emit=lambda i, p, g: emitted.__setitem__(i, self._coder.encode(g.detach().float())),
                        kill_signal=lambda: self.comm.kill_requested(self.cur_step), cur_step=self.cur_step)
                    comp_dur = time.time() - comp_start
                    if killed:
                        print("Worker: {}, Step: {} abandoned (PS already aggregated enough gradients)".format(
                            self.rank, self.cur_step))
                        self.comm.push(None, self.cur_step)
                        continue
                    encode_start = time.time()
                    msgs = [emitted[i] for i in range(len(emitted))]
                    msg_bytes = sum(Coding.wire_bytes(c) for c in msgs)
                    encode_dur = time.time() - encode_start
                else:
                    loss.backward()
                    comp_dur = time.time() - comp_start
                    encode_start = time.time()
                    msgs, msg_bytes = self._encode()
                    encode_dur = time.time() - encode_start

                comm_start = time.time()
                self._send_grads(msgs)
                comm_dur = time.time() - comm_start

                prec1, prec5 = accuracy(logits.detach(), y, topk=(1, 5))
                self.last_stats = dict(step=self.cur_step, loss=float(loss.item()), prec1=float(prec1.item()),
                                       prec5=float(prec5.item()), msg_bytes=msg_bytes, fetch=fetch_weight_duration)
                print(worker_line(self.rank, self.cur_step, num_epoch, batch_idx * self.batch_size, n_data,
                                  loss.item(), time.time() - iter_start, comp_dur, encode_dur, comm_dur,
                                  msg_bytes / (1024.0 ** 2), prec1.item(), prec5.item()))
                self._metrics.write(step=self.cur_step, epoch=num_epoch, loss=float(loss.item()),
                                    time=time.time() - iter_start, comp=comp_dur, encode=encode_dur, comm=comm_dur,
                                    fetch=fetch_weight_duration, msg_mb=msg_bytes / (1024.0 ** 2),
                                    prec1=float(prec1.item()), prec5=float(prec5.item()))
                if test_loader is not None and self.cur_step % self._eval_freq == 0:
                    self._evaluate_model(test_loader)
        # epochs exhausted before the PS stopped: keep answering until STOP
        while True:
            self.next_step = self.async_fetch_step()
            if self.next_step == STOP_STEP:
                break
            self.update_step()
            self.async_fetch_weights_bcast()
            self._send_grads([self._coder.encode(torch.zeros_like(p)) for p in self.network.parameters()])

    def async_fetch_step(self) -> int:
        return self.comm.recv_step()

    sync_fetch_step = async_fetch_step

    def update_step(self) -> bool:
        changed = self.cur_step != self.next_step
        self.cur_step = self.next_step
        return changed

    def async_fetch_weights_bcast(self):
        self.comm.fetch_params(self.flat_params)     # broadcast, or point-to-point in backup-worker rounds

    def _encode(self):
        msgs, nbytes = [], 0
        for p in self.network.parameters():
            coded = self._coder.encode(p.grad.detach().to(torch.float32))
            nbytes += Coding.wire_bytes(coded)
            msgs.append(coded)
        return msgs, nbytes

    def _send_grads(self, msgs):
        if self.rank == 1 and self._eval_freq and self.cur_step % self._eval_freq == 0:
            # checkpoint step: the first worker ships its BatchNorm running statistics with the gradients so the
            # PS's model_step_<N> carries TRAINED buffers (the PS never runs a forward pass; the reference's PS
            # checkpoints had untrained BN statistics, which is why it kept them off for ResNet, master:228-230)
            bufs = [b.detach().to("cpu") for b in self.network.buffers()]
            if bufs:
                msgs = list(msgs) + [{"__aux__": "buffers", "encode": False, "tensors": bufs}]
        self.comm.push(msgs, self.cur_step)

    def _generate_model_path(self):
        return ckpt.model_path(self._train_dir, self.cur_step)

    def _save_model(self, file_path=None):
        ckpt.save_model(self._train_dir, self.cur_step, self.network)

    @torch.no_grad()
    def _evaluate_model(self, test_loader):
        self.network.eval()
        test_loss, p1, p5, nb, n = 0.0, 0.0, 0.0, 0, 0
        for i, (data, target) in enumerate(test_loader):
            if self._eval_batches is not None and i >= self._eval_batches:
                break
            data, target = data.to(self.device), target.to(self.device)
            output = self.network(data)
            test_loss += F.cross_entropy(output, target, reduction="sum").item()
            a1, a5 = accuracy(output, target, topk=(1, 5))
            p1 += a1.item(); p5 += a5.item(); nb += 1; n += len(target)
        nb = max(nb, 1)
        print(test_line(self.cur_step, test_loss / max(n, 1), p1 / nb, p5 / nb))
        self.network.train()
        return {"loss": test_loss / max(n, 1), "prec1": p1 / nb, "prec5": p5 / nb}
