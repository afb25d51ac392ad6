"""Kernel timeline of one graph-replayed ShadowEngine step: which kernels run where, what is exposed after
backward.  python scripts/timeline_shadow.py [--code svd] [--groups 4] > gpurun_out/timeline.txt"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from atomo_b200.data import SyntheticImageDataset
from atomo_b200.models import build_model, input_shape
from atomo_b200.runtime.shadow_engine import ShadowEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--network", default="ResNet18")
    ap.add_argument("--code", default="svd")
    ap.add_argument("--groups", type=int, default=4)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--batch-size", type=int, default=128)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    eng = ShadowEngine(build_model(args.network, 10), 0, 1, code=args.code, svd_rank=3, lr=0.01, momentum=0.9,
                       use_graph=True, overlap=not args.no_overlap, groups=args.groups)
    x, y = SyntheticImageDataset(input_shape(args.network), 10, 4096).materialize(args.batch_size)
    eng.prepare(x.pin_memory(), y.pin_memory(), warmup=4)
    for _ in range(5):
        eng.train_step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            eng.train_step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if getattr(e, "device_time", 0) and "Memcpy" not in e.name and "Memset" not in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps by the wait_params kernel
    starts = [i for i, e in enumerate(evs) if "wait_params" in e.name]
    a, b = starts[1], starts[2]
    step = evs[a:b]
    t0 = step[0].time_range.start
    end = max(e.time_range.end for e in step)
    print("step span %.1f us, %d kernels" % (end - t0, len(step)))
    ours = [e for e in step if "atomo" in e.name]
    last_lib = max(e.time_range.end for e in step if "v2_" not in e.name)
    print("last non-v2 kernel ends at +%.1f us -> exposed tail %.1f us" % (last_lib - t0, end - last_lib))
    busy = sum(e.device_time for e in step if "v2_" not in e.name)
    print("sum of non-v2 kernel time %.1f us; sum of v2 kernel time %.1f us" % (busy, sum(e.device_time for e in step if "v2_" in e.name)))
    for e in step:
        if "v2_" in e.name:
            print("  +%8.1f .. +%8.1f  (%6.1f us)  %s" % (e.time_range.start - t0, e.time_range.end - t0, e.device_time,
                                                        e.name.split("(")[0].replace("atomo::v2::", "")))
    # gaps on the main chain: idle time between consecutive non-v2 kernels
    lib = [e for e in step if "v2_" not in e.name]
    gaps = sorted(((lib[i + 1].time_range.start - lib[i].time_range.end, lib[i].name[:50], lib[i + 1].name[:50])
                   for i in range(len(lib) - 1)), reverse=True)[:8]
    print("largest gaps between consecutive main-stream kernels:")
    for g, n1, n2 in gaps:
        print("  %6.1f us  after %s  before %s" % (g, n1, n2))
    eng.close()


if __name__ == "__main__":
    main()
