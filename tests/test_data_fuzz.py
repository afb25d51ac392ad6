"""Property tests of the data plumbing: shards are disjoint and equal-sized for any (n, workers), the persistent
loader wraps epochs without repeating inside one, checkpoint discovery ignores foreign files."""
import os

import pytest
import torch
import pytest as _pytest
_pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st

from atomo_b200.data import DataLoader, shard_indices
from atomo_b200.utils import checkpoint as ckpt


@settings(max_examples=60, deadline=None)
@given(n=st.integers(1, 3000), workers=st.integers(1, 17), seed=st.integers(0, 99), epoch=st.integers(0, 5))
def test_shards_are_disjoint_equal_and_seeded(n, workers, seed, epoch):
    shards = [shard_indices(n, w, workers, seed, epoch) for w in range(workers)]
    assert all(len(s) == n // workers for s in shards)
    flat = torch.cat(shards).tolist()
    assert len(set(flat)) == len(flat) and all(0 <= i < n for i in flat)
    again = shard_indices(n, workers - 1, workers, seed, epoch)
    assert torch.equal(again, shards[-1])


@settings(max_examples=20, deadline=None)
@given(n=st.integers(4, 200), bs=st.integers(1, 16), seed=st.integers(0, 50))
def test_persistent_loader_visits_every_sample_once_per_epoch(n, bs, seed):
    ds = torch.utils.data.TensorDataset(torch.arange(n).float().unsqueeze(1), torch.arange(n))
    if n // bs == 0:
        with pytest.raises(ValueError):
            DataLoader(ds, batch_size=bs, shuffle=True, drop_last=True, seed=seed, prefetch=0)
        return
    ld = DataLoader(ds, batch_size=bs, shuffle=True, drop_last=True, seed=seed, prefetch=0)
    per_epoch = n // bs
    for epoch in range(2):
        seen = []
        for _ in range(per_epoch):
            _, y = ld.next_batch()
            seen += y.tolist()
        assert len(set(seen)) == len(seen) == per_epoch * bs
    ld.next_batch()
    assert ld.epochs_completed == 2
    ld.close()


def test_latest_step_ignores_sidecars_temporaries_and_foreign_files(tmp_path):
    d = str(tmp_path) + "/"
    assert ckpt.latest_step(d) is None
    for name in ("model_step_4", "model_step_40", "model_step_40_optim", "model_step_100.tmp", "model_step_x",
                 "other_step_999", "model_step_7_optim"):
        open(os.path.join(d, name), "w").close()
    assert ckpt.latest_step(d) == 40
    assert ckpt.model_path(d, 40) == d + "model_step_40"          # plain concatenation, like the reference
    assert ckpt.latest_step(str(tmp_path / "run_")) is None       # a prefix-style train_dir is honoured
