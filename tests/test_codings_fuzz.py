"""Property tests over every registered coder: arbitrary tensor shapes (odd sizes, size-1 dimensions, all-zero and
tiny-magnitude gradients) must round-trip to the original shape with finite values, survive the wire format, and the
deterministic coders must be exact."""
import numpy as np
import pytest
import torch
import pytest as _pytest
_pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st

from atomo_b200 import codings
from atomo_b200.parallel import wire

shape_st = st.one_of(
    st.tuples(st.integers(1, 70)),
    st.tuples(st.integers(1, 40), st.integers(1, 40)),
    st.tuples(st.integers(1, 12), st.integers(1, 12), st.sampled_from([1, 3, 5]), st.sampled_from([1, 3, 5])),
    st.tuples(st.integers(1, 6), st.integers(1, 6), st.integers(1, 4)),
)
CODERS = {
    "sgd": dict(),
    "svd": dict(rank=3),
    "bsvd": dict(rank=3),
    "qsvd": dict(rank=2),
    "qsgd": dict(quantization_level=4, bucket_size=64),
    "terngrad": dict(bucket_size=64),
    "entrywise": dict(budget=0.3),
    "lossless": dict(),
}


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(shape=shape_st, name=st.sampled_from(sorted(CODERS)), kind=st.sampled_from(["randn", "zeros", "tiny", "spike"]),
       seed=st.integers(0, 10_000))
def test_every_coder_roundtrips_any_shape(shape, name, kind, seed):
    if name not in codings.available():
        pytest.skip(name)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    if kind == "zeros":
        x = torch.zeros(shape)
    elif kind == "tiny":
        x = x * 1e-20
    elif kind == "spike":
        x = torch.zeros(shape)
        x.view(-1)[seed % x.numel()] = 3.0
    coder = codings.build(name, **CODERS[name])
    code = coder.encode(x.clone())
    out = coder.decode(code)
    assert tuple(out.reshape(shape).shape) == tuple(shape)
    assert torch.isfinite(out).all()
    assert codings.Coding.wire_bytes(code) >= 0
    back = wire.unpack(wire.pack({"codes": [code]}))["codes"][0]
    assert torch.equal(coder.decode(back).reshape(shape), out.reshape(shape))
    if name in ("sgd", "lossless"):
        assert torch.equal(out.reshape(shape), x)
    if kind == "zeros":
        assert float(out.abs().max()) == 0.0


@settings(max_examples=25, deadline=None)
@given(shape=shape_st, seed=st.integers(0, 1000))
def test_spectral_coders_are_exact_at_full_budget(shape, seed):
    x = torch.randn(shape, generator=torch.Generator().manual_seed(seed))
    for name in ("svd", "bsvd"):
        coder = codings.build(name, rank=4096, random_sample=False)
        out = coder.decode(coder.encode(x)).reshape(shape)
        assert torch.allclose(out, x, atol=1e-4 * max(1.0, float(x.abs().max())))
