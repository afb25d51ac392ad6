"""Property tests of the v2 planner: the descriptor tables it emits are dereferenced by the CUDA kernels without any
bounds checks, so for ARBITRARY architectures every parameter element must be owned by exactly one PS tile, every
slot / Gram-partial / staging region must lie inside its buffer without overlapping another, and every tile must fit
the kernels' shared-memory and register-tile limits."""
import numpy as np
import pytest
import pytest as _pytest
_pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st

from atomo_b200.ops import plan2 as P


def check_plan(pl, owners, shapes):
    cover_w = np.zeros(pl.w_total, dtype=np.int32)
    cover_v = np.zeros(pl.v_total, dtype=np.int32)
    for (ui, a, b, o) in pl.ps_tiles:
        u = pl.units[ui]
        assert 0 <= o < owners and b > 0 and a >= 0
        if u.kind == P.KIND_VEC:
            assert a + b <= u.numel
            cover_v[u.w_off + a:u.w_off + a + b] += 1
        elif u.kind == P.KIND_DENSE16:
            assert a + b <= u.numel and b <= P.PS_TILE_ELEMS
            cover_w[u.w_off + a:u.w_off + a + b] += 1
        elif u.kind == P.KIND_SLAB:
            half = u.I // 2
            assert a % half == 0 and b % half == 0 and a + b <= u.rows
            n = (b // half) * u.K * u.I
            assert n <= P.PS_TILE_ELEMS
            e0 = u.w_off + (a // half) * u.K * u.I
            cover_w[e0:e0 + n] += 1
            assert P.owner_of_row(u, a, owners) == o
        else:
            assert u.kind == P.KIND_MAT and a + b <= u.rows
            assert b * u.cols <= P.PS_TILE_ELEMS and b <= P.PS_MAX_ROWS
            idx = (u.w_off + np.arange(a, a + b)[:, None] * u.rs + np.arange(u.cols)[None, :] * u.cs).reshape(-1)
            np.add.at(cover_w, idx, 1)
            assert P.owner_of_row(u, a, owners) == o
    assert len(pl.params) == len(shapes)
    for q, s in zip(pl.params, shapes):
        assert q.numel == int(np.prod(s))
        tgt = cover_w if q.is_w else cover_v
        assert (tgt[q.off:q.off + q.numel] == 1).all(), s
    # nothing outside the parameters is ever written (alignment gaps stay untouched)
    assert cover_w.sum() == sum(q.numel for q in pl.params if q.is_w)
    assert cover_v.sum() == sum(q.numel for q in pl.params if not q.is_w)

    # slots / Gram partials / staging copies: inside their buffers, pairwise disjoint
    def disjoint(regions, limit, what):
        regions = sorted(regions)
        for (lo, hi), nxt in zip(regions, regions[1:] + [(limit, limit)]):
            assert 0 <= lo < hi <= nxt[0], what
    coded = [u for u in pl.units if u.coded]
    ubits = 8 if pl.code == "qsvd" or any(u.ubits == 8 for u in coded) else 0
    disjoint([(u.slot_off, u.slot_off + P.slot_floats(u.rows, u.cols, u.rcap, u.ubits)) for u in coded],
             pl.arena_floats, "slots")
    disjoint([(u.gpart_off, u.gpart_off + u.n_enc * u.cols * u.cols) for u in coded], pl.gpart_floats, "gram partials")
    disjoint([(u.rs, u.rs + u.numel) for u in pl.units if u.kind == P.KIND_DENSE16], pl.stage_total, "staging")
    assert len({u.ts_index for u in coded}) == len(coded) == pl.n_coded
    for u in coded:
        assert 2 <= u.cols <= P.MAX_COLS and u.rcap % 4 == 0 and 4 <= u.rcap <= P.RCAP_MAX and u.rows >= 1
        assert u.budget >= 1 and (ubits == 0 or u.ubits == 8)
    # encode tiles: contiguous per group, complete per unit, inside the 36 KB tile buffer
    seen = 0
    for g, (t0, n) in enumerate(pl.enc_range):
        assert t0 == seen
        seen += n
        assert all(pl.units[ui].group == g for (ui, _, _, _) in pl.enc_tiles[t0:t0 + n])
    assert seen == len(pl.enc_tiles)
    for u in pl.units:
        tiles = pl.enc_tiles[u.enc_tile0:u.enc_tile0 + u.n_enc]
        if u.kind == P.KIND_SLAB:
            assert sum(t[2] for t in tiles) == u.rows // (u.I // 2)
            assert all(t[2] * u.K * (u.I // 2 + 4) * 4 <= 36 * 1024 for t in tiles)
        elif u.kind == P.KIND_MAT:
            assert sum(t[2] for t in tiles) == u.rows
            assert all(t[2] * ((u.cols + 3) // 4 * 4) * 4 <= 36 * 1024 for t in tiles)
    assert len(pl.units_bytes()) == P.UNIT_BYTES * len(pl.units)


conv = st.tuples(st.sampled_from([8, 16, 24, 64, 96, 130, 256]), st.sampled_from([1, 3, 8, 16, 32, 48, 64, 100, 256]),
                 st.sampled_from([1, 3, 5, 7])).map(lambda t: (t[0], t[1], t[2], t[2]))
linear = st.tuples(st.integers(2, 700), st.integers(2, 700))
vector = st.integers(1, 600).map(lambda n: (n,))
arch = st.lists(st.one_of(conv, linear, vector, conv), min_size=1, max_size=14)


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(shapes=arch, code=st.sampled_from(["svd", "qsvd", "sgd"]), rank=st.integers(1, 16),
       owners=st.integers(1, 8), groups=st.integers(1, 6), systematic=st.booleans())
def test_plans_of_random_architectures_are_memory_safe(shapes, code, rank, owners, groups, systematic):
    if not any(len(s) >= 2 for s in shapes):
        shapes = shapes + [(16, 16, 3, 3)]
    pl = P.build_plan2(shapes, code, rank, systematic, n_owners=owners, n_groups=groups)
    check_plan(pl, owners, shapes)


@pytest.mark.parametrize("net,ds,ncls", [("ResNet34", "Cifar10", 10), ("ResNet101", "ImageNet", 1000),
                                         ("DenseNet", "Cifar10", 10), ("AlexNet", "Cifar10", 10),
                                         ("VGG16", "Cifar10", 10), ("FC", "MNIST", 10)])
def test_plans_of_the_other_model_families(net, ds, ncls):
    from atomo_b200.models import build_model
    shapes = [tuple(p.shape) for p in build_model(net, ncls, ds).parameters()]
    for owners, code in ((1, "svd"), (8, "qsvd")):
        check_plan(P.build_plan2(shapes, code, 4, n_owners=owners, n_groups=5), owners, shapes)
