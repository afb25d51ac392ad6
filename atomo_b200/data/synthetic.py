"""Synthetic datasets of the shapes the reference trains on.

There is no network in the build/bench environment, so MNIST / CIFAR-10 /
CIFAR-100 / SVHN / "ImageNet" (``/root/reference/src/distributed_nn.py:93-207``)
are replaced by deterministic synthetic tensors of identical shape and class
count.  Samples are ``class_template[y] + noise`` so that a model can actually
fit them (loss decreases, accuracy rises) — useful for convergence tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch.utils.data import Dataset

SHAPES = {
    "mnist": ((1, 28, 28), 10, 60000, 10000),
    "cifar10": ((3, 32, 32), 10, 50000, 10000),
    "cifar100": ((3, 32, 32), 100, 50000, 10000),
    "svhn": ((3, 32, 32), 10, 73257, 26032),
    "imagenet": ((3, 224, 224), 1000, 1281167, 50000),
    # the reference's "ImageNet" branch is CIFAR-10 upscaled to 227x227 (launcher:175-207)
    "imagenet-ref": ((3, 227, 227), 10, 50000, 10000),
}


class SyntheticImageDataset(Dataset):
    """Deterministic, index-addressable synthetic image classification data.

    Items are generated on demand from ``(seed, index)`` so the dataset costs
    no memory regardless of ``length``; ``materialize()`` returns dense tensors
    for the GPU-resident fast path.
    """

    def __init__(self, shape: Tuple[int, int, int], num_classes: int, length: int,
                 seed: int = 0, noise: float = 0.5, train: bool = True):
        self.shape = tuple(shape)
        self.num_classes = int(num_classes)
        self.length = int(length)
        self.seed = int(seed)
        self.noise = float(noise)
        self.train = train
        g = torch.Generator().manual_seed(self.seed)
        # low-resolution class templates, upsampled: cheap and learnable
        c, h, w = self.shape
        lo = torch.randn(self.num_classes, c, max(h // 4, 1), max(w // 4, 1), generator=g)
        self.templates = torch.nn.functional.interpolate(lo, size=(h, w), mode="nearest")
        self.epochs_completed = 0

    def __len__(self):
        return self.length

    def _label(self, index: int) -> int:
        return (index * 2654435761 + self.seed + (0 if self.train else 7919)) % self.num_classes

    def __getitem__(self, index: int):
        y = self._label(index)
        g = torch.Generator().manual_seed((self.seed * 1000003 + index * 2 + (0 if self.train else 1)) & 0x7FFFFFFF)
        x = self.templates[y] + self.noise * torch.randn(self.shape, generator=g)
        return x, y

    def materialize(self, n: int = None, device="cpu"):
        n = min(n or self.length, self.length)
        idx = torch.arange(n)
        y = (idx * 2654435761 + self.seed + (0 if self.train else 7919)) % self.num_classes
        g = torch.Generator().manual_seed(self.seed + (0 if self.train else 1))
        x = self.templates[y] + self.noise * torch.randn((n,) + self.shape, generator=g)
        return x.to(device), y.to(device)


def synthetic_pair(name: str, seed: int = 0, train_len: int = None, test_len: int = None):
    key = name.lower()
    if key not in SHAPES:
        raise ValueError("unknown dataset %r" % name)
    shape, classes, ntrain, ntest = SHAPES[key]
    train = SyntheticImageDataset(shape, classes, train_len or ntrain, seed=seed, train=True)
    test = SyntheticImageDataset(shape, classes, test_len or ntest, seed=seed, train=False)
    return train, test
