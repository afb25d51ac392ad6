"""CPU tests of the v2 planner (ops/plan2.py): layouts, coverage, ownership, struct sizes."""
import struct

import pytest

from atomo_b200.ops import plan2 as P


def _shapes(net):
    from atomo_b200.models import build_model
    return [tuple(p.shape) for p in build_model(net, 10).parameters()]


def test_struct_sizes_match_the_cuda_side():
    assert P.UNIT_BYTES == 112 and P.TILE_BYTES == 16 and P.CTRL2_BYTES == 72
    c = P.pack_ctrl2(step=3, lr=0.5, seed=9, opt=P.OPT_ADAM)
    assert struct.unpack_from("<i", c, 0)[0] == 3 and struct.unpack_from("<Q", c, 32)[0] == 9


@pytest.mark.parametrize("net", ["ResNet18", "VGG11", "LeNet", "ResNet50"])
@pytest.mark.parametrize("owners", [1, 8])
def test_every_element_is_covered_exactly_once(net, owners):
    shapes = _shapes(net)
    pl = P.build_plan2(shapes, "svd", 3, n_owners=owners, n_groups=4)
    cover_w = [0] * pl.w_total
    cover_v = [0] * pl.v_total
    per_owner = [0] * owners
    for (ui, a, b, o) in pl.ps_tiles:
        u = pl.units[ui]
        assert 0 <= o < owners
        if u.kind == P.KIND_VEC:
            for e in range(u.w_off + a, u.w_off + a + b):
                cover_v[e] += 1
        elif u.kind == P.KIND_DENSE16:
            for e in range(u.w_off + a, u.w_off + a + b):
                cover_w[e] += 1
            per_owner[o] += b
        elif u.kind == P.KIND_SLAB:
            half = u.I // 2
            assert a % half == 0 and b % half == 0 and (b // half) * u.K * u.I <= P.PS_TILE_ELEMS
            e0 = u.w_off + (a // half) * u.K * u.I
            for e in range(e0, e0 + (b // half) * u.K * u.I):
                cover_w[e] += 1
            per_owner[o] += b * u.cols
            assert P.owner_of_row(u, a, owners) == o
        else:
            assert b * u.cols <= P.PS_TILE_ELEMS and b <= P.PS_MAX_ROWS
            for r in range(a, a + b):
                for c in range(u.cols):
                    cover_w[u.w_off + r * u.rs + c * u.cs] += 1
            per_owner[o] += b * u.cols
            assert P.owner_of_row(u, a, owners) == o
    for q in pl.params:
        rng = range(q.off, q.off + q.numel)
        tgt = cover_w if q.is_w else cover_v
        assert all(tgt[e] == 1 for e in rng), (net, q.shape)
    if owners > 1 and net != "LeNet":
        assert max(per_owner) < 1.3 * (sum(per_owner) / owners)      # the shards are balanced
    # encode tiles cover every coded unit exactly once, group ranges are contiguous and ordered
    seen = 0
    for g, (t0, n) in enumerate(pl.enc_range):
        assert t0 == seen
        seen += n
        for (ui, a, b, k) in pl.enc_tiles[t0:t0 + n]:
            assert pl.units[ui].group == g
    assert seen == len(pl.enc_tiles)
    for u in pl.units:
        tiles = pl.enc_tiles[u.enc_tile0:u.enc_tile0 + u.n_enc]
        assert [t[3] for t in tiles] == list(range(u.n_enc)) or u.kind == P.KIND_DENSE16
        if u.kind == P.KIND_SLAB:
            assert sum(t[2] for t in tiles) == u.rows // (u.I // 2)
            pitch = u.I // 2 + 4
            assert all(t[2] * u.K * pitch * 4 <= 36 * 1024 for t in tiles)
        elif u.kind == P.KIND_MAT:
            assert sum(t[2] for t in tiles) == u.rows
            assert all(t[2] * ((u.cols + 3) // 4 * 4) * 4 <= 36 * 1024 for t in tiles)


def test_groups_follow_backward_order_and_the_last_one_is_small():
    shapes = _shapes("ResNet18")
    g = P.default_groups(shapes, 5)
    w = [gi for gi, s in zip(g, shapes) if len(s) >= 2]
    assert w == sorted(w, reverse=True) and w[-1] == 0 and w[0] == 4
    assert w.count(4) == 1                      # the final group is the stem alone (dense: no eig on the tail)
    pl = P.build_plan2(shapes, "svd", 3, n_groups=5)
    share = [0] * pl.n_groups
    for q in pl.params:
        if q.is_w:
            share[q.group] += q.numel
    assert share[0] > 0.5 * sum(share) and share[-1] < 0.001 * sum(share) and share[-2] < 0.06 * sum(share)
    assert all(u.kind in (P.KIND_DENSE16, P.KIND_VEC) for u in pl.units if u.group == pl.n_groups - 1)
    # a BN vector rides in the group of the conv that precedes it
    for i, s in enumerate(shapes):
        if len(s) == 1 and i > 0:
            j = max(k for k in range(i) if len(shapes[k]) >= 2)
            assert g[i] == g[j]


def test_resnet18_unit_kinds():
    pl = P.build_plan2(_shapes("ResNet18"), "svd", 3)
    kinds = [u.kind for u in pl.units]
    assert kinds.count(P.KIND_SLAB) == 16          # every 3x3 conv except the 3-channel stem
    assert kinds.count(P.KIND_DENSE16) == 1        # the stem
    assert kinds.count(P.KIND_MAT) == 2 + 4 + 8 + 1  # 1x1 shortcuts in 32-column blocks + fc
    for u in pl.units:
        if u.kind == P.KIND_SLAB:
            assert u.cols == 18 and u.rcap == 8 and u.budget == 3.0
    dense = P.build_plan2(_shapes("ResNet18"), "sgd", 3)
    assert all(u.kind in (P.KIND_DENSE16, P.KIND_VEC) for u in dense.units) and dense.n_coded == 0


@pytest.mark.parametrize("net,ds,ncls,rank", [("ResNet18", "Cifar10", 10, 3), ("ResNet50", "ImageNet", 1000, 8),
                                              ("VGG11", "Cifar10", 10, 3)])
def test_sharded_owners_carry_equal_shares(net, ds, ncls, rank):
    """Sharded PS: tile j of a group goes to owner j % n_owners.  Every GPU must end up with the same share of the
    update work (elements reconstructed + updated + multicast), otherwise the slowest owner is the step's tail."""
    from atomo_b200.models import build_model
    shapes = [tuple(p.shape) for p in build_model(net, ncls, ds).parameters()]
    for owners in (2, 4, 8):
        pl = P.build_plan2(shapes, "svd", rank, False, n_owners=owners, n_groups=5)
        load = [0] * owners
        for (u, a, b, o) in pl.ps_tiles:
            un = pl.units[u]
            load[o] += b * un.cols if un.kind in (P.KIND_SLAB, P.KIND_MAT) else b
        assert sum(load) == sum(q.numel for q in pl.params)      # every element has exactly one owner
        assert max(load) <= 1.05 * sum(load) / owners, (net, owners, load)
