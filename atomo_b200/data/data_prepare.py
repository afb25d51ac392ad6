"""Dataset pre-fetch (parity: ``/root/reference/src/data/data_prepare.py:10-45``).

The reference downloads MNIST/CIFAR-10/CIFAR-100 once so that parallel ranks
do not race on the download.  Offline there is nothing to download: this tool
reports which datasets exist under ``--root`` and which will fall back to
synthetic data, and (``--materialize``) caches synthetic tensors as ``.pt``.
"""
import argparse
import os

import torch

from .datasets import _real_pair
from .synthetic import SHAPES, synthetic_pair


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default=".")
    ap.add_argument("--materialize", type=int, default=0, help="cache N synthetic samples per dataset")
    args = ap.parse_args(argv)
    for name in ("MNIST", "Cifar10", "Cifar100", "SVHN"):
        real = _real_pair(name, args.root)
        print("%-9s %s" % (name, "found on disk" if real is not None else "absent -> synthetic"))
        if real is None and args.materialize:
            train, _ = synthetic_pair(name.lower(), train_len=args.materialize)
            x, y = train.materialize(args.materialize)
            out = os.path.join(args.root, "%s_synthetic.pt" % name.lower())
            torch.save({"x": x, "y": y}, out)
            print("  wrote", out)


if __name__ == "__main__":
    main()
