"""Layer-by-layer ("split") models with overlapped gradient emission.

Parity: ``LeNetSplit`` (``/root/reference/src/model_ops/lenet.py:37-228``),
``FC_NN_Split`` (``fc_nn.py:33-152``) and ``ResNetSplit``
(``resnet_split.py:139-705``).  The reference hand-rolls, per model, a forward
that detaches between layers and five backward variants that ``Isend`` each
layer's gradient to the PS while the remaining backward is still running:

* ``backward`` / ``backward_normal`` — emit each layer's gradients as soon as
  they exist (resnet_split.py:259-456);
* ``backward_signal_kill`` — poll a kill signal between layers and abandon the
  step when the PS says this worker is a straggler (resnet_split.py:458-570,
  lenet.py:158-225); returns ``killed``;
* ``backward_timeout_kill`` — abandon the step after a time budget
  (resnet_split.py:572-684, ``timeout_decorator.timeout(10.5)``);
* ``backward_single`` — plain backward, no communication (resnet_split.py:686-705).

Here one generic :class:`SplitModel` does this for *any* sequence of stages: the
forward keeps per-stage ``(input, output)`` pairs with the graph cut between
stages, and ``backward*`` walks them in reverse, calling ``emit(param_index,
param, grad)`` the moment a stage's parameter gradients are ready.  The runtime
passes an ``emit`` that launches the encode+push kernels on a side stream, so
communication overlaps the rest of backward (on the fused GPU path the same
effect comes from ``register_post_accumulate_grad_hook``; this class preserves
the explicit, kill-able protocol).  Gradients are emitted last-layer first,
bias before weight, like the reference (lenet.py:122-154).
"""
from __future__ import annotations

import time
from typing import Callable, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .lenet import LeNet
from .fc_nn import FC_NN
from .resnet import BasicBlock, ResNet

Emit = Callable[[int, torch.nn.Parameter, torch.Tensor], None]


class _Fn(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


class SplitModel(nn.Module):
    def __init__(self, stages: Sequence, kill_threshold: Optional[float] = None):
        super().__init__()
        self.stages = nn.ModuleList([s if isinstance(s, nn.Module) else _Fn(s) for s in stages])
        self.kill_threshold = kill_threshold
        self.criterion = nn.CrossEntropyLoss()
        self._inputs: List[torch.Tensor] = []
        self._outputs: List[torch.Tensor] = []
        # global parameter index of each stage's parameters, parameters() order
        self._param_index = {}
        for i, p in enumerate(self.parameters()):
            self._param_index[id(p)] = i

    # ------------------------------------------------------------------
    def forward(self, x):
        self._inputs, self._outputs = [], []
        for stage in self.stages:
            x = x.detach().requires_grad_(True) if x.dtype.is_floating_point else x
            self._inputs.append(x)
            x = stage(x)
            self._outputs.append(x)
        return x

    def _walk(self, loss: torch.Tensor, emit: Optional[Emit], should_stop: Callable[[], bool]) -> bool:
        """Reverse walk; returns True when the step was abandoned."""
        grad_out = None
        n = len(self.stages)
        for i in range(n - 1, -1, -1):
            if should_stop():
                return True
            out = self._outputs[i]
            if i == n - 1:
                loss.backward()
            else:
                if grad_out is None:
                    break
                out.backward(grad_out)
            inp = self._inputs[i]
            grad_out = inp.grad if isinstance(inp, torch.Tensor) and inp.requires_grad else None
            if emit is not None:
                params = [p for p in self.stages[i].parameters() if p.grad is not None]
                for p in reversed(params):  # bias before weight (lenet.py:122-154)
                    emit(self._param_index[id(p)], p, p.grad)
        return False

    # -- the reference's five variants ----------------------------------
    def backward(self, loss, emit: Optional[Emit] = None, cur_step: int = 0):
        self._walk(loss, emit, lambda: False)

    backward_normal = backward

    def backward_single(self, loss):
        self._walk(loss, None, lambda: False)

    def backward_signal_kill(self, loss, emit: Optional[Emit] = None,
                             kill_signal: Optional[Callable[[], bool]] = None, cur_step: int = 0) -> bool:
        killed = self._walk(loss, emit, kill_signal or (lambda: False))
        return killed

    def backward_timeout_kill(self, loss, emit: Optional[Emit] = None,
                              timeout_s: Optional[float] = None, cur_step: int = 0) -> bool:
        budget = timeout_s if timeout_s is not None else (self.kill_threshold or 10.5)
        t0 = time.monotonic()
        return self._walk(loss, emit, lambda: (time.monotonic() - t0) > budget)

    def name(self):
        return "split"


# ----------------------------------------------------------------------
def LeNetSplit(num_classes: int = 10) -> SplitModel:
    base = LeNet(num_classes)
    return SplitModel([
        base.conv1, lambda x: F.relu(F.max_pool2d(x, 2, 2)),
        base.conv2, lambda x: F.relu(F.max_pool2d(x, 2, 2)).reshape(-1, 800),
        base.fc1, base.fc2,
    ])


def FC_NN_Split(num_classes: int = 10) -> SplitModel:
    base = FC_NN(num_classes)
    return SplitModel([
        lambda x: x.reshape(x.size(0), -1), base.fc1, base.relu,
        base.fc2, nn.ReLU(), base.fc3, base.sigmoid,
    ])


def _resnet_stages(net: ResNet):
    stages = [net.conv1, net.bn1, nn.ReLU()]
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        stages.extend(list(layer))  # one stage per residual block
    stages.append(lambda x: F.avg_pool2d(x, 4).flatten(1))
    stages.append(net.linear)
    return stages


def ResNetSplit18(kill_threshold: Optional[float] = None, num_classes: int = 10) -> SplitModel:
    return SplitModel(_resnet_stages(ResNet(BasicBlock, [2, 2, 2, 2], num_classes)), kill_threshold)


def ResNetSplit34(kill_threshold: Optional[float] = None, num_classes: int = 10) -> SplitModel:
    return SplitModel(_resnet_stages(ResNet(BasicBlock, [3, 4, 6, 3], num_classes)), kill_threshold)
