"""DenseNet-BC for 32x32 inputs
(parity: ``/root/reference/src/model_ops/densenet.py:18-116``;
the runtime builds DenseNet-BC-190-40, ``sync_replicas_master_nn.py:156-158``).

Divergence: ``forward`` returns raw logits.  The reference applies
``log_softmax`` and then trains with ``CrossEntropyLoss`` (double softmax,
SURVEY.md 2.9); ``log_softmax_output=True`` restores that behaviour.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Bottleneck(nn.Module):
    def __init__(self, nChannels, growthRate):
        super().__init__()
        inter = 4 * growthRate
        self.bn1 = nn.BatchNorm2d(nChannels)
        self.conv1 = nn.Conv2d(nChannels, inter, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(inter)
        self.conv2 = nn.Conv2d(inter, growthRate, 3, padding=1, bias=False)

    def forward(self, x):
        out = self.conv1(F.relu(self.bn1(x)))
        out = self.conv2(F.relu(self.bn2(out)))
        return torch.cat((x, out), 1)


class SingleLayer(nn.Module):
    def __init__(self, nChannels, growthRate):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(nChannels)
        self.conv1 = nn.Conv2d(nChannels, growthRate, 3, padding=1, bias=False)

    def forward(self, x):
        return torch.cat((x, self.conv1(F.relu(self.bn1(x)))), 1)


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
        n = (depth - 4) // 3
        if bottleneck:
            n //= 2
        ch = 2 * growthRate
        self.conv1 = nn.Conv2d(3, ch, 3, padding=1, bias=False)
        self.dense1 = self._make_dense(ch, growthRate, n, bottleneck)
        ch += n * growthRate
        out = int(math.floor(ch * reduction))
        self.trans1 = Transition(ch, out)
        ch = out
        self.dense2 = self._make_dense(ch, growthRate, n, bottleneck)
        ch += n * growthRate
        out = int(math.floor(ch * reduction))
        self.trans2 = Transition(ch, out)
        ch = out
        self.dense3 = self._make_dense(ch, growthRate, n, bottleneck)
        ch += n * growthRate
        self.bn1 = nn.BatchNorm2d(ch)
        self.fc = nn.Linear(ch, nClasses)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                k = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / k))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.bias.data.zero_()

    @staticmethod
    def _make_dense(ch, growthRate, n, bottleneck):
        layers = []
        for _ in range(int(n)):
            layers.append(Bottleneck(ch, growthRate) if bottleneck else SingleLayer(ch, growthRate))
            ch += growthRate
        return nn.Sequential(*layers)

    def forward(self, x):
        out = self.conv1(x)
        out = self.trans1(self.dense1(out))
        out = self.trans2(self.dense2(out))
        out = self.dense3(out)
        out = F.avg_pool2d(F.relu(self.bn1(out)), 8).flatten(1)
        out = self.fc(out)
        return F.log_softmax(out, dim=1) if self.log_softmax_output else out
