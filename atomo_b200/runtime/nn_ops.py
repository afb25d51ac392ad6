"""Base trainer, accuracy, and the single-process training loop.

Parity: ``/root/reference/src/nn_ops.py`` — ``accuracy`` (86-99),
``NN_Trainer`` with ``build_model`` / ``train_and_validate`` / ``validate``
(101-189) and the ``svd_encode`` indicator probe (66-82).  The duplicated
``_resize_to_2d`` / ``_sample_svd`` copies of the reference live once in
``atomo_b200.codings``.
"""
from __future__ import annotations

import time
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import codings
from ..codings.svd import resize_to_2d
from ..codings.utils import l1_indicator, nuclear_indicator
from ..models import build_model


def accuracy(output: torch.Tensor, target: torch.Tensor, topk: Tuple[int, ...] = (1,)):
    """Precision@k in percent (nn_ops.py:86-99); clamps k to the class count."""
    with torch.no_grad():
        maxk = min(max(topk), output.size(1))
        batch_size = target.size(0)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
        res = []
        for k in topk:
            k = min(k, maxk)
            res.append(correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / batch_size))
        return res


def svd_encode(grad: torch.Tensor, step: int = 0, verbose: bool = True):
    """Indicator probe (nn_ops.py:66-82): which atom family suits this gradient."""
    mat = resize_to_2d(grad).to(torch.float32)
    s = torch.linalg.svdvals(mat)
    nuc, l1 = nuclear_indicator(mat, s), l1_indicator(mat)
    if verbose:
        print("Step: {}, Nuclear Indicator: {}, L1 Indicator: {}".format(step, nuc, l1))
    return nuc, l1


class NN_Trainer:
    """Single-process trainer (also the base class of the PS/worker/evaluator)."""

    def __init__(self, **kwargs):
        self.batch_size = kwargs.get("batch_size", 128)
        self.lr = kwargs.get("learning_rate", 0.01)
        self.max_epochs = kwargs.get("max_epochs", 1)
        self.momentum = kwargs.get("momentum", 0.5)
        self.network_config = kwargs.get("network", "LeNet")
        self.dataset = kwargs.get("dataset", "")
        self.device = torch.device(kwargs.get("device", "cpu"))
        self.fetch_indicator = kwargs.get("fetch_indicator", False)
        self.log_interval = kwargs.get("log_interval", 10)
        # optional: pass every gradient through a coder (encode -> decode) before the optimizer step: the effect of
        # a sparsifier on training, without a cluster (one worker, so no averaging of independent draws)
        self.coder = kwargs.get("coder", None)

    def build_model(self, num_classes: int = 10):
        self.network = build_model(self.network_config, num_classes, self.dataset).to(self.device)
        self.optimizer = torch.optim.SGD(self.network.parameters(), lr=self.lr, momentum=self.momentum)
        self.criterion = nn.CrossEntropyLoss()
        return self

    def train_and_validate(self, train_loader, test_loader, max_steps: Optional[int] = None):
        step = 0
        for epoch in range(self.max_epochs):
            self.network.train()
            for batch_idx, (x, y) in enumerate(train_loader):
                t0 = time.time()
                x, y = x.to(self.device), y.to(self.device)
                self.optimizer.zero_grad()
                logits = self.network(x)
                loss = self.criterion(logits, y)
                loss.backward()
                if self.fetch_indicator:
                    for p in self.network.parameters():
                        svd_encode(p.grad, step)
                if self.coder is not None:
                    for p in self.network.parameters():
                        g = self.coder.decode(self.coder.encode(p.grad.detach().float()))
                        p.grad = g.reshape(p.shape).to(p.grad.device, p.grad.dtype)
                self.optimizer.step()
                prec1, prec5 = accuracy(logits, y, topk=(1, 5))
                step += 1
                if step % self.log_interval == 0:
                    print("Train Epoch: {} [{}/{} ({:.0f}%)]  Loss: {:.4f}, Time: {:.4f}, Prec@1: {:.4f}, Prec@5: {:.4f}".format(
                        epoch, batch_idx * len(x), len(train_loader.dataset),
                        100.0 * batch_idx / max(len(train_loader), 1), loss.item(), time.time() - t0,
                        prec1.item(), prec5.item()))
                if max_steps is not None and step >= max_steps:
                    return self.validate(test_loader)
            self.validate(test_loader)
        return self.validate(test_loader)

    @torch.no_grad()
    def validate(self, test_loader, max_batches: Optional[int] = None):
        self.network.eval()
        loss_sum, p1, p5, n, nb = 0.0, 0.0, 0.0, 0, 0
        for i, (x, y) in enumerate(test_loader):
            if max_batches is not None and i >= max_batches:
                break
            x, y = x.to(self.device), y.to(self.device)
            out = self.network(x)
            loss_sum += F.cross_entropy(out, y, reduction="sum").item()
            a1, a5 = accuracy(out, y, topk=(1, 5))
            p1 += a1.item(); p5 += a5.item(); n += len(y); nb += 1
        nb = max(nb, 1)
        res = {"loss": loss_sum / max(n, 1), "prec1": p1 / nb, "prec5": p5 / nb}
        print("Test set: Average loss: {:.4f}, Prec@1: {} Prec@5: {}".format(res["loss"], res["prec1"], res["prec5"]))
        return res
