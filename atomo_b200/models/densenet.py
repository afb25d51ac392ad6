"""DenseNet(-BC) for 32x32 inputs (capability parity: ``/root/reference/src/model_ops/densenet.py``; the
runtime builds DenseNet-BC-190-40, ``sync_replicas_master_nn.py:156-158``).

One ``_DenseLayer`` covers both the bottleneck (BN-ReLU-1x1 -> BN-ReLU-3x3) and the single (BN-ReLU-3x3)
variants; module names (``dense1..3``, ``trans1..2``, ``bn1/conv1/bn2/conv2``, ``fc``) and the parameter order
match the reference.  Divergence: ``forward`` returns raw logits — the reference applies ``log_softmax`` and then
trains with ``CrossEntropyLoss`` (a double softmax); ``log_softmax_output=True`` restores that.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _DenseLayer(nn.Module):
    def __init__(self, in_ch: int, growth: int, bottleneck: bool):
        super().__init__()
        self.bottleneck = bottleneck
        self.bn1 = nn.BatchNorm2d(in_ch)
        if bottleneck:
            mid = 4 * growth
            self.conv1 = nn.Conv2d(in_ch, mid, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(mid)
            self.conv2 = nn.Conv2d(mid, growth, 3, padding=1, bias=False)
        else:
            self.conv1 = nn.Conv2d(in_ch, growth, 3, padding=1, bias=False)

    def forward(self, x):
        new = self.conv1(F.relu(self.bn1(x)))
        if self.bottleneck:
            new = self.conv2(F.relu(self.bn2(new)))
        return torch.cat((x, new), 1)


def Bottleneck(nChannels, growthRate):
    return _DenseLayer(nChannels, growthRate, True)


def SingleLayer(nChannels, growthRate):
    return _DenseLayer(nChannels, growthRate, False)


class Transition(nn.Module):
    def __init__(self, nChannels, nOutChannels):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(nChannels)
        self.conv1 = nn.Conv2d(nChannels, nOutChannels, 1, bias=False)

    def forward(self, x):
        return F.avg_pool2d(self.conv1(F.relu(self.bn1(x))), 2)


class DenseNet(nn.Module):
    def __init__(self, growthRate=12, depth=100, reduction=0.5, nClasses=10, bottleneck=True,
                 log_softmax_output=False):
        super().__init__()
        self.log_softmax_output = log_softmax_output
        per_block = (depth - 4) // 3 // (2 if bottleneck else 1)
        ch = 2 * growthRate
        self.conv1 = nn.Conv2d(3, ch, 3, padding=1, bias=False)
        for stage in (1, 2, 3):
            layers = [_DenseLayer(ch + i * growthRate, growthRate, bottleneck) for i in range(per_block)]
            setattr(self, "dense%d" % stage, nn.Sequential(*layers))
            ch += per_block * growthRate
            if stage < 3:
                squeezed = int(math.floor(ch * reduction))
                setattr(self, "trans%d" % stage, Transition(ch, squeezed))
                ch = squeezed
        self.bn1 = nn.BatchNorm2d(ch)
        self.fc = nn.Linear(ch, nClasses)
        self._init_weights()

    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan))
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)

    def forward(self, x):
        out = self.conv1(x)
        out = self.trans1(self.dense1(out))
        out = self.trans2(self.dense2(out))
        out = self.dense3(out)
        out = self.fc(F.avg_pool2d(F.relu(self.bn1(out)), 8).flatten(1))
        return F.log_softmax(out, dim=1) if self.log_softmax_output else out
