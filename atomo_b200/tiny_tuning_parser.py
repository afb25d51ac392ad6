"""LR-sweep log parser (parity: ``/root/reference/src/tiny_tuning_parser.py:4-27``):
reads ``<tuning-dir>/<lr>`` files holding the grepped ``Step: N`` worker log
lines of one trial, regex-parses the loss of every worker and prints the mean.
The worker log-line format (``utils/logging.py``) is the contract."""
import argparse
import os
import re

LINE_RE = re.compile(
    r"Worker: (\d+), Step: (\d+), Epoch: (\d+) \[(\d+)/(\d+) \((\d+)%\)\], Loss: ([-\d.naninfe+]+), "
    r"Time Cost: ([\d.]+), Comp: ([\d.]+), Encode:\s+([\d.]+), Comm:\s+([\d.]+), Msg\(MB\):\s+([\d.]+), "
    r"Prec@1:\s+([\d.]+), Prec@5:\s+([\d.]+)")


def parse_line(line: str):
    m = LINE_RE.search(line)
    if not m:
        return None
    g = m.groups()
    return {"worker": int(g[0]), "step": int(g[1]), "epoch": int(g[2]), "loss": float(g[6]),
            "time": float(g[7]), "comp": float(g[8]), "encode": float(g[9]), "comm": float(g[10]),
            "msg_mb": float(g[11]), "prec1": float(g[12]), "prec5": float(g[13])}


def mean_loss(path: str, num_workers: int = None, step: int = None):
    losses = []
    with open(path) as f:
        for line in f:
            rec = parse_line(line)
            if rec is None or (step is not None and rec["step"] != step):
                continue
            losses.append(rec["loss"])
    if not losses:
        return None
    if num_workers:
        losses = losses[-num_workers:]
    return sum(losses) / len(losses)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--tuning-dir", type=str, default="tune/")
    ap.add_argument("--tuning-lr", type=str, default="0.01")
    ap.add_argument("--num-workers", type=int, default=16)
    ap.add_argument("--step", type=int, default=0)
    args = ap.parse_args(argv)
    path = os.path.join(args.tuning_dir, str(args.tuning_lr))
    loss = mean_loss(path, args.num_workers, args.step or None)
    print("Learning rate: {}, Avged loss: {}".format(args.tuning_lr, loss))
    return loss


if __name__ == "__main__":
    main()
