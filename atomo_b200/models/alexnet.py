"""AlexNet, spec-driven (capability parity: ``/root/reference/src/model_ops/alexnet.py`` — the
224/227-input layout with a ``256*6*6`` classifier; ``features.N`` / ``classifier.N`` module indices and the
parameter order are those of the reference so checkpoints and per-tensor coders line up)."""
import torch.nn as nn

# (out_channels, kernel, stride, padding, max-pool after?)
_FEATURES = ((64, 11, 4, 2, True), (192, 5, 1, 2, True), (384, 3, 1, 1, False), (256, 3, 1, 1, False),
             (256, 3, 1, 1, True))
_HIDDEN = (4096, 4096)


class AlexNet(nn.Module):
    def __init__(self, num_classes=1000):
        super().__init__()
        feats, ch = [], 3
        for out, k, s, p, pool in _FEATURES:
            feats += [nn.Conv2d(ch, out, kernel_size=k, stride=s, padding=p), nn.ReLU(inplace=True)]
            if pool:
                feats.append(nn.MaxPool2d(kernel_size=3, stride=2))
            ch = out
        self.features = nn.Sequential(*feats)
        self.flat_features = ch * 6 * 6
        head, width = [], self.flat_features
        for h in _HIDDEN:
            head += [nn.Dropout(), nn.Linear(width, h), nn.ReLU(inplace=True)]
            width = h
        head.append(nn.Linear(width, num_classes))
        self.classifier = nn.Sequential(*head)

    def forward(self, x):
        return self.classifier(self.features(x).reshape(x.size(0), self.flat_features))


def alexnet(pretrained=False, **kwargs):
    """``pretrained`` exists for signature parity only: there is no network to download weights from."""
    if pretrained:
        raise RuntimeError("pretrained weights are not available offline")
    return AlexNet(**kwargs)
