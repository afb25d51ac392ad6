"""Per-phase and per-kernel breakdown of one fused training step (run on the GPU box).

    python scripts/profile_step.py [--dtype bf16] [--channels-last] [--batch-size 128] [--out gpurun_out/profile.txt]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from atomo_b200.data import SyntheticImageDataset
from atomo_b200.models import build_model, input_shape
from atomo_b200.runtime.engine import FusedEngine


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--network", default="ResNet18")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--channels-last", action="store_true")
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--code", default="svd")
    ap.add_argument("--out", default="gpurun_out/profile.txt")
    ap.add_argument("--kernels", action="store_true", help="also dump a torch.profiler kernel table")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = build_model(args.network, 10, "Cifar10")
    eng = FusedEngine(model, 0, 1, code=args.code, svd_rank=3, lr=0.01, momentum=0.9, dtype=args.dtype,
                      channels_last=args.channels_last, use_graph=True)
    x, y = SyntheticImageDataset(input_shape(args.network, "Cifar10"), 10, 4096).materialize(args.batch_size)
    eng.prepare(x.pin_memory(), y.pin_memory(), warmup=3)
    lines = ["config: %s" % vars(args)]
    lines.append("graph step           : %.3f ms" % timed(lambda: eng.train_step()))

    def fb():
        eng.flat_grads.zero_()
        if eng.bn_arena is not None:
            eng.bn_arena.zero_()
        eng._forward_backward()
    lines.append("eager fwd+bwd        : %.3f ms" % timed(fb))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fb()
    lines.append("graphed fwd+bwd      : %.3f ms" % timed(g.replay))
    lines.append("eager encode+push    : %.3f ms" % timed(eng._encode_push))
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        eng._encode_push()
    lines.append("graphed encode+push  : %.3f ms" % timed(g2.replay))
    C, pl = eng.C, eng.plan
    if args.code == "svd":
        lines.append("  gram               : %.3f ms" % timed(lambda: C.gram(eng.flat_grads, eng.t_layers, eng.t_enc_tiles, len(pl.enc_tiles), eng.gpart)))
        arena0 = eng.heap.region_ptr("arena", 0)
        lines.append("  eig_sample         : %.3f ms" % timed(lambda: C.eig_sample(eng.t_layers, eng.t_ts_layers, eng.gpart, eng.vsel, eng.selcount, eng.sigma, arena0, pl.arena_floats, eng.ctrl, None, 3, True, False, False, 0)))
        lines.append("  project_push       : %.3f ms" % timed(lambda: C.project_push(eng.flat_grads, eng.t_layers, eng.t_enc_tiles, len(pl.enc_tiles), eng.vsel, eng.selcount, arena0, pl.arena_floats, eng.ps_push_flags, eng.ctrl, 0, True)))
    # ps_update alone: the flags are already >= step, so it does not block
    eng.ctrl_i32[0] = 1
    lines.append("ps_update            : %.3f ms" % timed(eng._ps_update))
    lines.append("bytes: params %.1f MB, factors/worker(cap) %.1f MB, dense %.2f MB" % (
        pl.total_elems * 4 / 2**20, pl.factor_bytes_per_worker() / 2**20, pl.dense_bytes() / 2**20))
    if args.kernels:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                eng.flat_grads.zero_()
                if eng.bn_arena is not None:
                    eng.bn_arena.zero_()
                eng._forward_backward()
                eng._encode_push()
                eng._ps_update()
            torch.cuda.synchronize()
        lines.append(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main()
