"""Per-launch device time of every BN kernel inside one real ResNet-18 forward/backward (eager, torch.profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from scripts.probe_bf16_leaf import make, fwd_bwd

def main():
    torch.cuda.set_device(0)
    torch.backends.cudnn.benchmark = True
    model, arena, x, y = make(True)
    model.train()
    fn = lambda: fwd_bwd(model, arena, x, y, True)
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if "bn_" in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    tot = {}
    for e in evs:
        nm = e.name.split("(")[0].replace("atomo::", "")
        d = e.device_time if hasattr(e, "device_time") else e.cuda_time
        tot[nm] = tot.get(nm, 0) + d
        print("%-28s %7.1f us" % (nm, d))
    print(tot)

if __name__ == "__main__":
    main()
