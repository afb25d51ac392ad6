"""VGG 11/13/16/19 (+BN) for 32x32 inputs
(parity: ``/root/reference/src/model_ops/vgg.py:15-107``): conv features,
512-512-512-classes classifier with Dropout, He-normal conv init, cfgs A/B/D/E.
Every variant takes ``num_classes`` (only ``vgg11_bn`` does in the reference).
"""
import math

import torch.nn as nn

cfg = {
    "A": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "B": [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "D": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
    "E": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M",
          512, 512, 512, 512, "M"],
}


class VGG(nn.Module):
    def __init__(self, features, num_classes=10):
        super().__init__()
        self.features = features
        self.classifier = nn.Sequential(
            nn.Dropout(), nn.Linear(512, 512), nn.ReLU(True),
            nn.Dropout(), nn.Linear(512, 512), nn.ReLU(True),
            nn.Linear(512, num_classes),
        )
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
                m.bias.data.zero_()

    def forward(self, x):
        return self.classifier(self.features(x).flatten(1))


def make_layers(c, batch_norm=False):
    layers, in_ch = [], 3
    for v in c:
        if v == "M":
            layers.append(nn.MaxPool2d(2, 2))
        else:
            layers.append(nn.Conv2d(in_ch, v, 3, padding=1))
            if batch_norm:
                layers.append(nn.BatchNorm2d(v))
            layers.append(nn.ReLU(inplace=True))
            in_ch = v
    return nn.Sequential(*layers)


def _vgg(key, bn, num_classes):
    return VGG(make_layers(cfg[key], batch_norm=bn), num_classes=num_classes)


def vgg11(num_classes=10): return _vgg("A", False, num_classes)
def vgg11_bn(num_classes=10): return _vgg("A", True, num_classes)
def vgg13(num_classes=10): return _vgg("B", False, num_classes)
def vgg13_bn(num_classes=10): return _vgg("B", True, num_classes)
def vgg16(num_classes=10): return _vgg("D", False, num_classes)
def vgg16_bn(num_classes=10): return _vgg("D", True, num_classes)
def vgg19(num_classes=10): return _vgg("E", False, num_classes)
def vgg19_bn(num_classes=10): return _vgg("E", True, num_classes)
