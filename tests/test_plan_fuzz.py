"""Property tests of the fp32-flat engine's planner (ops/plan.py), same purpose as tests/test_plan2_fuzz.py: the
kernels of csrc/{svd,ps,ext}_kernels.cu trust these tables blindly."""
import numpy as np
import pytest as _pytest
_pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st

from atomo_b200.ops import plan as P


def check(pl, shapes):
    cover = np.zeros(pl.total_elems, dtype=np.int32)
    for (li, a, b, c0) in pl.ps_tiles:
        l = pl.layers[li]
        if l.route == P.ROUTE_SVD_TS:
            assert 0 <= a and a + b <= l.rows and 0 < b <= l.ps_rows
            idx = (l.off + np.arange(a, a + b)[:, None] * l.row_stride + np.arange(l.cols)[None, :] * l.col_stride)
            np.add.at(cover, idx.reshape(-1), 1)
        elif l.route == P.ROUTE_LOWRANK_EXT:
            nc = min(P.TS_MAX_COLS, l.cols - c0)
            assert 0 <= a and a + b <= l.rows and 0 <= c0 < l.cols and nc > 0
            idx = (l.off + np.arange(a, a + b)[:, None] * l.row_stride + np.arange(c0, c0 + nc)[None, :] * l.col_stride)
            np.add.at(cover, idx.reshape(-1), 1)
        else:
            assert 4 * a + b <= l.numel and 0 < b <= P.PS_DENSE_ELEMS
            cover[l.off + 4 * a:l.off + 4 * a + b] += 1
    dense = np.zeros(pl.total_elems, dtype=np.int32)
    for (li, a, b, _) in pl.dense_tiles:
        l = pl.layers[li]
        dense[l.off + 4 * a:l.off + 4 * a + b] += 1
    for l, s in zip(pl.layers, shapes):
        assert l.numel == int(np.prod(s)) and l.rows * l.cols == l.numel
        assert (cover[l.off:l.off + l.numel] == 1).all() and (dense[l.off:l.off + l.numel] == 1).all(), s
    total = sum(l.numel for l in pl.layers)
    assert cover.sum() == total and dense.sum() == total

    def disjoint(regions, limit):
        regions = sorted(regions)
        for (lo, hi), nxt in zip(regions, regions[1:] + [(limit, limit)]):
            assert 0 <= lo < hi <= nxt[0]
    coded = [l for l in pl.layers if l.route != P.ROUTE_DENSE]
    disjoint([(l.slot_off, l.slot_off + P.slot_floats(l.rows, l.cols, l.rcap)) for l in coded], pl.arena_floats)
    ts = [l for l in coded if l.route == P.ROUTE_SVD_TS]
    disjoint([(l.gpart_off, l.gpart_off + l.ntiles * l.cols * l.cols) for l in ts], pl.gpart_floats)
    for l in ts:
        assert 2 <= l.cols <= P.TS_MAX_COLS and l.rows >= l.cols and l.rcap % 4 == 0 and l.rcap <= P.RCAP_MAX
        tiles = pl.enc_tiles[l.tile0:l.tile0 + l.ntiles]
        assert all(t[0] == l.index for t in tiles) and sum(t[2] for t in tiles) == l.rows
        assert [t[1] for t in tiles] == list(np.cumsum([0] + [t[2] for t in tiles[:-1]]))
    assert len(pl.layers_bytes()) == P.LAYER_BYTES * len(pl.layers)


conv = st.tuples(st.sampled_from([6, 16, 64, 100, 256]), st.sampled_from([1, 3, 16, 20, 64, 128]),
                 st.sampled_from([1, 3, 5])).map(lambda t: (t[0], t[1], t[2], t[2]))
linear = st.tuples(st.integers(2, 600), st.integers(2, 600))
vector = st.integers(1, 700).map(lambda n: (n,))


@settings(max_examples=100, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(shapes=st.lists(st.one_of(conv, linear, vector), min_size=1, max_size=12),
       code=st.sampled_from(["svd", "sgd", "qsgd", "entrywise"]), rank=st.integers(0, 12), systematic=st.booleans(),
       subspace=st.booleans(), dense_vectors=st.booleans())
def test_fp32_plans_of_random_architectures_are_memory_safe(shapes, code, rank, systematic, subspace, dense_vectors):
    pl = P.build_plan(shapes, code, rank, systematic, dense_vectors=dense_vectors, subspace=subspace, ext_min_numel=4096)
    check(pl, shapes)
