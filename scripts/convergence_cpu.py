#!/usr/bin/env python
"""Statistical-efficiency check on CPU: does the estimator the bf16 engine applies (block-spectral, `bsvd`) train like
the reference's whole-tensor spectral coder (`svd`) at the same rank?

Simulates the synchronous PS in one process (W workers, each with its own batches; every gradient goes through
coder.encode -> coder.decode; the PS averages and applies momentum SGD), so the only difference between the arms is
the coder.  Same seeds, same batches, same initial weights for every arm.

    python scripts/convergence_cpu.py --network LeNet --steps 150 --workers 4 --out docs/experiments/convergence_lenet.md
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from atomo_b200 import codings                      # noqa: E402
from atomo_b200.data import SyntheticImageDataset   # noqa: E402
from atomo_b200.models import build_model, input_shape  # noqa: E402


def run_arm(name, coder, args, shape, ncls):
    torch.manual_seed(args.seed)
    model = build_model(args.network, ncls, args.dataset)
    opt = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum)
    train = SyntheticImageDataset(shape, ncls, 1 << 20, seed=args.seed, noise=args.noise)
    test = SyntheticImageDataset(shape, ncls, 4096, seed=args.seed, noise=args.noise, train=False)
    xt, yt = test.materialize(args.test_len)
    gen = torch.Generator().manual_seed(args.seed + 1)
    if hasattr(coder, "generator"):
        coder.generator = gen
    W, B = args.workers, args.batch_size
    xs, ys = train.materialize(args.steps * W * B)
    losses, sent, t0 = [], 0, time.time()
    for step in range(args.steps):
        agg = [torch.zeros_like(p) for p in model.parameters()]
        step_loss = 0.0
        for w in range(W):
            lo = (step * W + w) * B
            model.zero_grad(set_to_none=True)
            loss = F.cross_entropy(model(xs[lo:lo + B]), ys[lo:lo + B])
            loss.backward()
            step_loss += float(loss.detach()) / W
            for a, p in zip(agg, model.parameters()):
                code = coder.encode(p.grad.detach())
                sent += codings.Coding.wire_bytes(code)
                a += coder.decode(code).reshape(p.shape) / W
        for a, p in zip(agg, model.parameters()):
            p.grad = a
        opt.step()
        losses.append(step_loss)
    model.eval()
    with torch.no_grad():
        out = torch.cat([model(xt[i:i + 256]) for i in range(0, len(xt), 256)])
    k = max(args.steps // 10, 1)
    return {"arm": name, "final_train_loss": sum(losses[-k:]) / k, "loss_at_25pct": sum(losses[args.steps // 4:args.steps // 4 + k]) / k,
            "test_loss": float(F.cross_entropy(out, yt)), "test_prec1": float((out.argmax(1) == yt).float().mean()) * 100,
            "MB_per_worker_step": sent / args.steps / W / 2 ** 20, "seconds": time.time() - t0, "losses": losses}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--network", default="LeNet")
    ap.add_argument("--dataset", default="")
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--rank", type=int, default=3)
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--momentum", type=float, default=0.9)
    ap.add_argument("--noise", type=float, default=2.0, help="per-pixel noise of the synthetic task (templates have unit variance)")
    ap.add_argument("--test-len", type=int, default=2048)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    if not args.dataset:
        args.dataset = "MNIST" if args.network in ("LeNet", "FC") else "Cifar10"
    ncls = 10
    shape = input_shape(args.network, args.dataset)
    arms = [("sgd (dense)", codings.build("sgd")),
            ("svd r=%d (reference estimator, whole tensor)" % args.rank, codings.build("svd", rank=args.rank)),
            ("bsvd r=%d (bf16 engine's estimator: slabs + 32-column blocks)" % args.rank, codings.build("bsvd", rank=args.rank)),
            ("qsvd r=%d" % args.rank, codings.build("qsvd", rank=args.rank))]
    rows = [run_arm(n, c, args, shape, ncls) for n, c in arms]
    lines = ["# CPU convergence check: %s, %d workers x batch %d, %d steps, lr %g, momentum %g, synthetic noise %g"
             % (args.network, args.workers, args.batch_size, args.steps, args.lr, args.momentum, args.noise), "",
             "Produced by `scripts/convergence_cpu.py` (one process simulating the synchronous PS; identical seeds, batches and",
             "initial weights in every arm; every gradient passes through `encode -> decode`).", "",
             "| coder | train loss at 25 % | final train loss (last 10 %) | test loss | test prec@1 | MB / worker / step |",
             "|---|---|---|---|---|---|"]
    for r in rows:
        lines.append("| %s | %.4f | %.4f | %.4f | %.2f | %.3f |" % (r["arm"], r["loss_at_25pct"], r["final_train_loss"],
                                                                   r["test_loss"], r["test_prec1"], r["MB_per_worker_step"]))
    text = "\n".join(lines) + "\n"
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
        with open(os.path.splitext(args.out)[0] + ".json", "w") as f:
            json.dump([{k: v for k, v in r.items()} for r in rows], f)


if __name__ == "__main__":
    main()
