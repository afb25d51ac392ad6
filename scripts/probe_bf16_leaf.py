"""GPU probe (run under gpurun): what does a ResNet-18 step cost when the conv / linear weights are bf16
channels_last LEAF tensors (no autocast weight casts, no NHWC weight copies, no fp32 grad casts, grads stolen
by AccumulateGrad instead of added)?  Also checks the mechanics the overlapped engine relies on:
grad pointers are stable under CUDA-graph replay, post-accumulate-grad hooks can fork a side stream
inside a capture.  Prints one JSON line per experiment.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

from atomo_b200.data import SyntheticImageDataset
from atomo_b200.models import build_model
from atomo_b200.ops.fused_bn import enable_fused_bn


def timed(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def make(leaf_bf16: bool, batch=128):
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = build_model("ResNet18", 10).to(dev).to(memory_format=torch.channels_last)
    n, arena = enable_fused_bn(model, True, arena_device=dev)
    if leaf_bf16:
        for p in model.parameters():
            if p.dim() >= 2:
                mf = torch.channels_last if p.dim() == 4 else torch.contiguous_format
                p.data = p.data.to(torch.bfloat16).contiguous(memory_format=mf)
    x, y = SyntheticImageDataset((3, 32, 32), 10, 4096).materialize(batch)
    x = x.to(dev).contiguous(memory_format=torch.channels_last)
    y = y.to(dev)
    return model, arena, x, y


def fwd_bwd(model, arena, x, y, none_grads):
    if none_grads:
        for p in model.parameters():
            p.grad = None
    if arena is not None:
        arena.zero_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = nn.functional.cross_entropy(model(x).float(), y)
    loss.backward()
    return loss


def kernel_table(fn, tag):
    from torch.profiler import profile, ProfilerActivity
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    rows = {}
    for e in prof.key_averages():
        if e.device_type.name == "CUDA" or getattr(e, "self_device_time_total", 0) > 0:
            t = getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)
            rows[e.key[:90]] = (e.count, t)
    total = sum(t for _, t in rows.values())
    top = sorted(rows.items(), key=lambda kv: -kv[1][1])[:28]
    with open("gpurun_out/probe_kernels_%s.txt" % tag, "w") as f:
        f.write("total device us %.1f, kernels %d\n" % (total, sum(c for c, _ in rows.values())))
        for k, (c, t) in top:
            f.write("%8.1f us  x%-4d %s\n" % (t, c, k))
    return total, sum(c for c, _ in rows.values())


def main():
    torch.cuda.set_device(0)
    torch.backends.cudnn.benchmark = True
    os.makedirs("gpurun_out", exist_ok=True)
    out = {}
    for leaf in (False, True):
        tag = "leaf_bf16" if leaf else "fp32_autocast"
        model, arena, x, y = make(leaf)
        model.train()
        none = leaf
        if not leaf:
            for p in model.parameters():
                p.grad = torch.zeros_like(p)

            def fn():
                for p in model.parameters():
                    p.grad.zero_()
                return fwd_bwd(model, arena, x, y, False)
        else:
            def fn():
                return fwd_bwd(model, arena, x, y, True)
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        info = {"mode": tag}
        if leaf:
            stolen, dtypes, cl = 0, set(), 0
            for p in model.parameters():
                if p.dim() >= 2:
                    dtypes.add(str(p.grad.dtype))
                    stolen += int(p.grad.stride() == p.stride())
                    cl += int(p.dim() == 4 and p.grad.is_contiguous(memory_format=torch.channels_last))
            info.update(grad_dtypes=sorted(dtypes), grad_layout_match=stolen, grad_channels_last=cl)
        tot, nk = kernel_table(fn, tag)
        info.update(eager_device_us=round(tot, 1), kernels_per_step=nk)
        info["eager_ms"] = round(timed(fn), 4)
        # graph
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        info["graph_ms"] = round(timed(g.replay), 4)
        if leaf:
            ptrs = [p.grad.data_ptr() for p in model.parameters()]
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            info["grad_ptrs_stable"] = ptrs == [p.grad.data_ptr() for p in model.parameters()]
            gn = float(torch.stack([p.grad.float().norm() for p in model.parameters()]).norm())
            info["grad_norm_after_replay"] = round(gn, 5)
        out[tag] = info
        print(json.dumps(info), flush=True)
        del g

    # ---- hooks forking a side stream inside a capture -------------------------------------------------
    model, arena, x, y = make(True)
    model.train()
    params = [p for p in model.parameters()]
    side = torch.cuda.Stream()
    evs = {}
    marks = torch.zeros(len(params), device="cuda")
    done_ev = torch.cuda.Event()
    hook_order = []

    def make_hook(i):
        def hook(p):
            hook_order.append(i)
            ev = evs.setdefault(i, torch.cuda.Event())
            ev.record(torch.cuda.current_stream())
            side.wait_event(ev)
            with torch.cuda.stream(side):
                marks[i] = p.grad.float().abs().sum()      # stands in for gram/eig/project of this layer
        return hook
    heavy = [i for i, p in enumerate(params) if p.dim() == 4][::5]
    for i in heavy:
        params[i].register_post_accumulate_grad_hook(make_hook(i))

    def fn2():
        fwd_bwd(model, arena, x, y, True)
        done_ev.record(side)
        torch.cuda.current_stream().wait_event(done_ev)
    for _ in range(3):
        hook_order.clear()
        fn2()
    torch.cuda.synchronize()
    eager_marks = marks.clone()
    res = {"mode": "hook_fork", "hooks": len(heavy), "hook_order_first": hook_order[:4]}
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn2()
        marks.zero_()
        g.replay()
        torch.cuda.synchronize()
        res["capture_ok"] = True
        res["marks_match"] = bool(torch.allclose(marks[heavy], eager_marks[heavy], rtol=2e-2))
        res["graph_ms"] = round(timed(g.replay), 4)
    except Exception as e:  # noqa
        res["capture_ok"] = False
        res["error"] = repr(e)[:300]
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
