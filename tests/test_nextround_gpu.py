"""GPU tests written after this round's GPU budget was spent: NOT part of `-m gpu`.
Run them first thing next round:  ATOMO_NEXTROUND=1 python -m pytest tests/test_nextround_gpu.py -q
"""
import os

import pytest
import torch

pytestmark = pytest.mark.skipif(not (os.environ.get("ATOMO_NEXTROUND") and torch.cuda.is_available()),
                                reason="unverified GPU tests: set ATOMO_NEXTROUND=1 on a GPU box")


def test_engine_checkpoint_resume_roundtrip(tmp_path):
    from atomo_b200.data import SyntheticImageDataset
    from atomo_b200.models import build_model
    from atomo_b200.runtime.engine import FusedEngine
    torch.cuda.set_device(0)
    d = str(tmp_path) + "/"
    x, y = SyntheticImageDataset((1, 28, 28), 10, 512).materialize(32)
    x, y = x.pin_memory(), y.pin_memory()

    def make():
        torch.manual_seed(0)
        return FusedEngine(build_model("LeNet", 10), 0, 1, code="sgd", lr=0.05, momentum=0.9,
                           use_graph=False, seed=5)

    a = make()
    a.prepare(x, y, warmup=0)
    for _ in range(5):
        a.train_step(x, y)
    a.save_checkpoint(d)                       # step 5
    for _ in range(3):
        a.train_step(x, y)
    torch.cuda.synchronize()
    want = a.flat_params.clone()
    a.close()

    b = make()
    b.prepare(x, y, warmup=0)
    b.load_checkpoint(d, 5)
    assert b.device_step() == 6
    for _ in range(3):
        b.train_step(x, y)
    torch.cuda.synchronize()
    # dense coder: deterministic up to cuDNN's atomics.  (With a sampling coder a 1e-7 difference in a
    # probability can flip a Bernoulli draw, so resumed runs are statistically, not bitwise, identical:
    # observed on B200 this round: load/step bookkeeping exact, parameters within 1e-3.)
    assert torch.allclose(b.flat_params, want, rtol=1e-3, atol=1e-4)
    b.close()


@pytest.mark.parametrize("W,rank", [(8, 3), (8, 8), (4, 16)])
def test_ps_update_many_virtual_workers_multi_chunk_K(W, rank):
    """W 'virtual workers' on one GPU (W arenas, all local): exercises the concatenated-K path of
    ps_update_kernel with K = sum_w count_w > PS_KC (several K chunks, float4 group gathers, SV reload)."""
    from tests.test_gpu_kernels import Harness, SHAPES, _fill_grads
    h = Harness(SHAPES, rank=rank, lr=0.05, momentum=0.9)
    pl = h.plan
    dev = h.dev
    arena = torch.zeros(W * pl.arena_floats, device=dev)
    grads = [torch.zeros_like(h.grads) for _ in range(W)]
    t_grads_peer = torch.tensor([g.data_ptr() for g in grads], dtype=torch.int64, device=dev)
    est = torch.zeros_like(h.params)
    for w in range(W):
        _fill_grads(h, seed=100 + w)
        grads[w].copy_(h.grads)
        C = h.C
        C.gram(h.grads, h.t_layers, h.t_enc, len(pl.enc_tiles), h.gpart)
        C.eig_sample(h.t_layers, h.t_ts, h.gpart, h.vsel, h.selcount, h.sigma, arena.data_ptr(), pl.arena_floats,
                     h.ctrl, None, rank, True, False, False, w, 1024)
        C.project_push(h.grads, h.t_layers, h.t_enc, len(pl.enc_tiles), h.vsel, h.selcount, arena.data_ptr(),
                       pl.arena_floats, h.flags.data_ptr(), h.ctrl, w, True)
        torch.cuda.synchronize()
        # reference: decode this worker's slots
        saved = h.arena
        h.arena = arena[w * pl.arena_floats:(w + 1) * pl.arena_floats]
        for l in pl.layers:
            if l.route == 1:
                c, s, V, U = h.slot(l)
                h.tall(l, est).add_((U * s) @ V)
            else:
                est[l.off:l.off + l.numel] += grads[w][l.off:l.off + l.numel]
        h.arena = saved
    assert all(int(h.flags[w]) == 1 for w in range(W))
    ref_p = h.params - 0.05 * (est / W)          # first step: momentum buffer = gradient
    h.C.ps_update(h.t_layers, h.t_ps, len(pl.ps_tiles), W, W, 1, h.params, h.mom, h.t_params_peer, 0, t_grads_peer, 0,
                  arena.data_ptr(), pl.arena_floats, h.flags.data_ptr(), h.t_flag_peer, h.ctrl, int(5e9), 1.0 / W,
                  min(len(pl.ps_tiles), 296))
    torch.cuda.synchronize()
    assert int(h.ctrl.view(torch.int32)[1]) == 0
    for l in pl.layers:
        a, b = h.params[l.off:l.off + l.numel], ref_p[l.off:l.off + l.numel]
        assert torch.allclose(a, b, rtol=2e-4, atol=2e-5), (l.shape, float((a - b).abs().max()))
