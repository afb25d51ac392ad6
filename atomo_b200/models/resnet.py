"""ResNet family (parity: ``/root/reference/src/model_ops/resnet.py:14-127``).

CIFAR-style stem (3x3 conv, no max-pool, ``avg_pool2d(4)``) like the reference
(resnet.py:74-111) plus an ImageNet stem (7x7/2 conv + 3x3/2 max-pool +
adaptive average pool) for BASELINE config 4 (ResNet-50 ImageNet-shaped).  All
depths construct and run — the reference's ResNet34/50/101/152 constructors
crash; parameter ordering for ResNet18 is identical to the
reference so the per-tensor coder sees the same 62 tensors.
"""
import torch.nn as nn
import torch.nn.functional as F

from ..ops.fused_bn import BNAct


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, stride, 1, bias=False)
        self.bn1 = BNAct(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = BNAct(planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_planes, self.expansion * planes, 1, stride, bias=False),
                BNAct(self.expansion * planes),
            )

    def forward(self, x):
        out = self.bn1(self.conv1(x), relu=True)
        # bn2 + residual add + ReLU in one (optionally fused) op
        return self.bn2(self.conv2(out), residual=self.shortcut(x), relu=True)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 1, bias=False)
        self.bn1 = BNAct(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = BNAct(planes)
        self.conv3 = nn.Conv2d(planes, self.expansion * planes, 1, bias=False)
        self.bn3 = BNAct(self.expansion * planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_planes, self.expansion * planes, 1, stride, bias=False),
                BNAct(self.expansion * planes),
            )

    def forward(self, x):
        out = self.bn1(self.conv1(x), relu=True)
        out = self.bn2(self.conv2(out), relu=True)
        return self.bn3(self.conv3(out), residual=self.shortcut(x), relu=True)


class ResNet(nn.Module):
    def __init__(self, block, num_blocks, num_classes=10, imagenet_stem=False):
        super().__init__()
        self.in_planes = 64
        self.imagenet_stem = imagenet_stem
        if imagenet_stem:
            self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        else:
            self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.bn1 = BNAct(64)
        self.layer1 = self._make_layer(block, 64, num_blocks[0], 1)
        self.layer2 = self._make_layer(block, 128, num_blocks[1], 2)
        self.layer3 = self._make_layer(block, 256, num_blocks[2], 2)
        self.layer4 = self._make_layer(block, 512, num_blocks[3], 2)
        self.linear = nn.Linear(512 * block.expansion, num_classes)

    def _make_layer(self, block, planes, n, stride):
        layers = []
        for s in [stride] + [1] * (n - 1):
            layers.append(block(self.in_planes, planes, s))
            self.in_planes = planes * block.expansion
        return nn.Sequential(*layers)

    def forward(self, x):
        out = self.bn1(self.conv1(x), relu=True)
        if self.imagenet_stem:
            out = F.max_pool2d(out, 3, 2, 1)
        out = self.layer4(self.layer3(self.layer2(self.layer1(out))))
        if self.imagenet_stem:
            out = F.adaptive_avg_pool2d(out, 1)
        else:
            out = F.avg_pool2d(out, 4)
        return self.linear(out.flatten(1))


def ResNet18(num_classes=10, **kw):
    return ResNet(BasicBlock, [2, 2, 2, 2], num_classes, **kw)


def ResNet34(num_classes=10, **kw):
    return ResNet(BasicBlock, [3, 4, 6, 3], num_classes, **kw)


def ResNet50(num_classes=10, **kw):
    return ResNet(Bottleneck, [3, 4, 6, 3], num_classes, **kw)


def ResNet101(num_classes=10, **kw):
    return ResNet(Bottleneck, [3, 4, 23, 3], num_classes, **kw)


def ResNet152(num_classes=10, **kw):
    return ResNet(Bottleneck, [3, 8, 36, 3], num_classes, **kw)
