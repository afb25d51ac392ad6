"""Kernel table of one eager ShadowEngine step (torch.profiler device times).  python scripts/profile_shadow.py [--overlap]"""
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
    ap.add_argument("--overlap", action="store_true")
    ap.add_argument("--groups", type=int, default=4)
    ap.add_argument("--batch-size", type=int, default=128)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    eng = ShadowEngine(build_model(args.network, 10), 0, 1, code=args.code, svd_rank=3, lr=0.01, momentum=0.9,
                       use_graph=False, overlap=args.overlap, groups=args.groups)
    x, y = SyntheticImageDataset(input_shape(args.network), 10, 4096).materialize(args.batch_size)
    eng.prepare(x.pin_memory(), y.pin_memory(), warmup=4)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.train_step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type.name == "CUDA" or "v2_" in e.name]
    ours = [e for e in prof.events() if "v2_" in e.name]
    ours.sort(key=lambda e: e.time_range.start)
    t0 = min([e.time_range.start for e in prof.events()] or [0])
    print("config", vars(args), "units", len(eng.plan.units), "enc tiles", [c for _, c in eng.plan.enc_range])
    for e in ours:
        d = e.device_time if hasattr(e, "device_time") else e.cuda_time
        print("%8.1f us  +%8.1f  %s" % (d, e.time_range.start - t0, e.name.split("(")[0]))
    tot = {}
    for e in prof.events():
        d = getattr(e, "device_time", 0) or 0
        if d and "Memcpy" not in e.name:
            tot[e.name[:60]] = tot.get(e.name[:60], 0) + d
    print("total device us: %.1f" % sum(tot.values()))
    eng.close()

if __name__ == "__main__":
    main()
