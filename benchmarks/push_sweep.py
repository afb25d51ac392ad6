#!/usr/bin/env python
"""Gradient-push bandwidth sweep (BASELINE.json config 5).

For rank budgets r in {1,2,4,8,16} (spectral) and entry-wise budgets s in {1%,5%,25%} this times, on the
device and as the max over ranks, the communication half of a step on ResNet-18-shaped gradients:

  ours  : encode + NVLink peer push (all workers) + PS reconstruct/aggregate + SGD + NVLS multicast
  nccl  : the same number of wire bytes per worker sent to rank 0 with torch.distributed (NCCL send/recv)
          followed by a dist.broadcast of the dense parameters — the reference's transport, minus its
          host staging, SVD and decode (i.e. a lower bound on the NCCL path's push+broadcast time).

Reported: effective GB/s = dense gradient bytes represented (W x 42.6 MB) / time, and wire GB/s.
Launch with torchrun (or plain python for 1 GPU).  One JSON line per configuration from rank 0.

NOTE: one engine (one symmetric heap + NVLS binding) per process: run one configuration per launch
(``--ranks 4 --budgets ""``); ``scripts/push_sweep.sh`` loops over the grid.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--network", default="ResNet18")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--ranks", default="1,2,4,8,16")
    ap.add_argument("--budgets", default="0.01,0.05,0.25")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from atomo_b200.models import build_model
    from atomo_b200.runtime.engine import FusedEngine

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    configs = [("svd", int(r), 0.0) for r in args.ranks.split(",") if r] + \
              [("entrywise", 0, float(b)) for b in args.budgets.split(",") if b]
    for code, r, budget in configs:
        torch.manual_seed(0)
        eng = FusedEngine(build_model(args.network, 10), rank, world, code=code, svd_rank=max(r, 1), lr=0.01,
                          momentum=0.9, use_graph=False, entry_budget=budget or 0.05, seed=3, timeout_s=60.0,
                          subspace="auto")
        g = torch.Generator(device="cuda").manual_seed(rank + 1)
        eng.flat_grads.copy_(torch.randn(eng.flat_grads.shape, device=dev, generator=g) * 1e-2)

        def comm_step():
            eng.C.wait_params(eng.local_param_flag, eng.ctrl, eng.timeout_ticks, 0)
            if eng.is_worker:
                eng._encode_push()
            if eng.is_ps:
                eng._ps_update()
            eng.C.advance_step(eng.ctrl)

        for _ in range(5):
            comm_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            comm_step()
        e1.record()
        barrier()
        ms = maxr(e0.elapsed_time(e1)) / args.iters
        dense_bytes = eng.plan.total_elems * 4
        if code == "svd":
            # measured message: headers + s + V + only the float4 groups of U that carry atoms
            cnt = eng.selcount.float().mean().item() if eng.plan.enc_tiles else 0
            wire = sum(4 * (4 + l.rcap + l.rcap * l.cols + l.rows * 4 * ((min(l.rcap, max(int(round(cnt)), 1)) + 3) // 4))
                       for l in eng.plan.layers if l.route != 0) + eng.plan.dense_bytes()
        else:
            wire = int(8 * eng._entry_expected(eng.plan.total_elems))
        # NCCL transport of the same bytes: W sends to rank 0 + dense parameter broadcast
        nccl_ms = None
        if world > 1:
            payload = torch.empty(wire // 4, dtype=torch.float32, device=dev)
            params = eng.flat_params.clone()

            def nccl_step():
                if rank == 0:
                    bufs = [torch.empty_like(payload) for _ in range(world - 1)]
                    reqs = [dist.irecv(b, src=w) for b, w in zip(bufs, range(1, world))]
                    for q in reqs:
                        q.wait()
                else:
                    dist.send(payload, dst=0)
                dist.broadcast(params, src=0)

            for _ in range(3):
                nccl_step()
            barrier()
            e0.record()
            for _ in range(args.iters):
                nccl_step()
            e1.record()
            barrier()
            nccl_ms = maxr(e0.elapsed_time(e1)) / args.iters
        if rank == 0:
            W = eng.W
            print(json.dumps({
                "bench": "push_sweep", "n_gpus": world, "workers": W, "code": code, "rank_budget": r,
                "entry_budget": budget, "ms_encode_push_decode_sgd_bcast": round(ms, 4),
                "wire_MB_per_worker": round(wire / 2 ** 20, 3),
                "effective_GBps": round(W * dense_bytes / (ms * 1e-3) / 1e9, 1),
                "wire_GBps_into_ps": round(W * wire / (ms * 1e-3) / 1e9, 2),
                "nccl_transport_only_ms": None if nccl_ms is None else round(nccl_ms, 4),
                "nvls_multicast": eng.heap.has_multicast, "device_error": eng.error_code()}))
        eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
