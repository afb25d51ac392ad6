"""Latency of the fused Gram+eig encode launch for single units (what sits on the critical path of the last
backward group), versus the Jacobi sweep cap.  Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_v2 import H2


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for shape in [(64, 64, 3, 3), (128, 64, 1, 1), (512, 512, 3, 3), (48, 16, 5, 5)]:
    for warm, sw in [(False, 12), (True, 3), (True, 2), (True, 1)]:
        h = H2([shape], rank=3, warm=warm, max_sweeps=sw if warm else 0)
        h.fill(0, 1)
        pl = h.plan
        gptr = torch.tensor([t.data_ptr() for t in h.wgrads[0]], dtype=torch.int64, device=h.dev)
        t0, nt = pl.enc_range[0]

        def enc():
            h.C.v2_encode(h.t_units.data_ptr(), h.t_enc.data_ptr(), t0, nt, gptr.data_ptr(), h.gpart.data_ptr(),
                          h.counters.data_ptr(), h.vsel.data_ptr(), h.selcount.data_ptr(), h.sigma.data_ptr(),
                          h.t_arena_peer.data_ptr(), 1, pl.arena_floats, h.stage[0].data_ptr(), h.ctrl.data_ptr(), 0,
                          h.vprev.data_ptr() if h.vprev is not None else 0, h.max_sweeps, True, False, False, 0, False, 0, 0, 0)
        us = timed(enc)
        print("%-18s units %d cols %s tiles %3d  warm=%-5s sweeps<=%-2d  encode launch %.1f us" %
              (shape, pl.n_coded, [u.cols for u in pl.units if u.coded][:1], nt, warm, sw, us))
