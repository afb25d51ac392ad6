#!/usr/bin/env python
"""Singular-value decay of a conv-layer gradient over training (parity: the reference's
``images/SVdecay.jpg`` — singular values of one 3x3-conv gradient, matricized to ``M x 18``, at data
passes 0 / 5 / 10: steep decay after rank 2-3, which is what makes a rank-3 budget work).

Prints (and optionally saves as CSV) the 18 singular values of the chosen layer's matricized gradient
every ``--every`` steps of a short local training run on synthetic CIFAR-shaped data.

    python tools/sv_decay.py --network ResNet18 --layer layer1.0.conv1.weight --steps 30 --every 10
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from atomo_b200.codings.svd import resize_to_2d
from atomo_b200.data import SyntheticImageDataset
from atomo_b200.models import build_model, input_shape


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--network", default="ResNet18")
    ap.add_argument("--layer", default="layer1.0.conv1.weight")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--every", type=int, default=10)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--csv", default="")
    args = ap.parse_args(argv)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    torch.manual_seed(0)
    model = build_model(args.network, 10).to(dev)
    param = dict(model.named_parameters())[args.layer]
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    ds = SyntheticImageDataset(input_shape(args.network), 10, 4096)
    x, y = ds.materialize(args.batch_size * 8, device=dev)
    rows = []
    for step in range(args.steps + 1):
        i = (step % 8) * args.batch_size
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x[i:i + args.batch_size]), y[i:i + args.batch_size]).backward()
        if step % args.every == 0:
            s = torch.linalg.svdvals(resize_to_2d(param.grad.detach().float()))
            s = (s / s[0]).cpu().tolist()
            rows.append((step, s))
            print("step %4d  %s" % (step, " ".join("%.3f" % v for v in s)))
        opt.step()
    if args.csv:
        with open(args.csv, "w") as f:
            for step, s in rows:
                f.write("%d,%s\n" % (step, ",".join("%.6f" % v for v in s)))
    return rows


if __name__ == "__main__":
    main()
