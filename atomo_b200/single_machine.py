"""Single-process trainer (parity: ``/root/reference/src/single_machine.py``:
argparse 29-56, main 185-256 -> ``NN_Trainer.train_and_validate``).  The
reference's file cannot even be imported (``from cifar10 import cifar10``,
line 25); this one runs on CPU or one GPU and can probe the spectral vs
entry-wise indicators per layer (``--fetch-indicator``)."""
import argparse

import torch

from .data import build_datasets
from .runtime.nn_ops import NN_Trainer
from .utils.flags import bool_flag


def main(argv=None):
    ap = argparse.ArgumentParser(description="atomo_b200 single-machine trainer")
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--test-batch-size", type=int, default=1000)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--max-steps", type=int, default=0)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--momentum", type=float, default=0.5)
    ap.add_argument("--no-cuda", action="store_true", default=False)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--log-interval", type=int, default=10)
    ap.add_argument("--network", type=str, default="LeNet")
    ap.add_argument("--dataset", type=str, default="MNIST")
    ap.add_argument("--synthetic", type=bool_flag, default=None)
    ap.add_argument("--data-root", type=str, default=".")
    ap.add_argument("--train-len", type=int, default=0)
    ap.add_argument("--test-len", type=int, default=0)
    ap.add_argument("--fetch-indicator", type=bool_flag, default=False)
    ap.add_argument("--code", type=str, default="",
                    help="pass every gradient through this coder (svd | bsvd | qsvd | qsgd | terngrad | entrywise) "
                         "before the optimizer step: what a sparsifier does to training, without a cluster")
    ap.add_argument("--svd-rank", type=int, default=3)
    ap.add_argument("--quantization-level", type=int, default=4)
    ap.add_argument("--entry-budget", type=float, default=0.05)
    args = ap.parse_args(argv)

    torch.manual_seed(args.seed)
    device = "cuda" if (torch.cuda.is_available() and not args.no_cuda) else "cpu"
    train, test, ncls = build_datasets(args.dataset, args.data_root, args.synthetic, args.seed,
                                       args.train_len or None, args.test_len or None)
    train_loader = torch.utils.data.DataLoader(train, batch_size=args.batch_size, shuffle=True)
    test_loader = torch.utils.data.DataLoader(test, batch_size=args.test_batch_size, shuffle=False)
    coder = None
    if args.code and args.code not in ("sgd", "dense"):
        from .runtime.master import build_coder
        coder = build_coder({"code": args.code, "svd_rank": args.svd_rank, "entry_budget": args.entry_budget,
                             "quantization_level": args.quantization_level}, worker_side=True)
    trainer = NN_Trainer(coder=coder, batch_size=args.batch_size, learning_rate=args.lr, max_epochs=args.epochs,
                         momentum=args.momentum, network=args.network, dataset=args.dataset, device=device,
                         fetch_indicator=args.fetch_indicator, log_interval=args.log_interval)
    trainer.build_model(num_classes=ncls)
    return trainer.train_and_validate(train_loader, test_loader, max_steps=args.max_steps or None)


if __name__ == "__main__":
    main()
