"""DataLoader with a persistent iterator and ``next_batch()``.

Parity: ``/root/reference/src/data_loader_ops/my_data_loader.py`` — a fork of
the torch-0.3 ``DataLoader`` whose only additions are a persistent iterator and
``next_batch()`` (310-319), backed by worker processes (37-53) and a
pin-memory thread (56-75).  Here the modern ``torch.utils.data.DataLoader``
provides the workers; this class adds the persistent-iterator API, per-epoch
seeded shuffling, and a background *pinned-memory staging thread* that keeps
``prefetch`` batches ready in page-locked buffers so the training loop's
``H2D`` copy is a single async ``cudaMemcpyAsync`` per tensor.
"""
from __future__ import annotations

import queue
import threading
from typing import Iterator, Optional

import torch
from torch.utils.data import DataLoader as _TorchLoader, Dataset


class DataLoader:
    def __init__(self, dataset: Dataset, batch_size: int = 1, shuffle: bool = False, num_workers: int = 0,
                 pin_memory: bool = False, drop_last: bool = False, seed: int = 0, prefetch: int = 2):
        self.dataset = dataset
        self.batch_size = batch_size
        self.pin_memory = pin_memory and torch.cuda.is_available()
        self._gen = torch.Generator().manual_seed(seed)
        self._loader = _TorchLoader(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers,
                                    drop_last=drop_last, generator=self._gen, pin_memory=False,
                                    persistent_workers=num_workers > 0)
        if len(self._loader) == 0:
            raise ValueError("DataLoader: %d samples give no batch of %d with drop_last=%s (a worker's shard is "
                             "smaller than --batch-size?)" % (len(dataset), batch_size, drop_last))
        self._iter: Optional[Iterator] = None
        self.epochs_completed = 0
        self._prefetch = max(int(prefetch), 0)
        self._q: Optional[queue.Queue] = None
        self._thread: Optional[threading.Thread] = None
        self._stop = threading.Event()

    def __len__(self):
        return len(self._loader)

    def __iter__(self):
        return iter(self._loader)

    # -- persistent iteration ------------------------------------------
    def _raw_next(self):
        if self._iter is None:
            self._iter = iter(self._loader)
        try:
            return next(self._iter)
        except StopIteration:
            self.epochs_completed += 1
            self._iter = iter(self._loader)
            return next(self._iter)

    def _pin(self, batch):
        if not self.pin_memory:
            return batch
        return tuple(t.pin_memory() if isinstance(t, torch.Tensor) else t for t in batch)

    def _stage_loop(self):
        while not self._stop.is_set():
            item = self._pin(self._raw_next())
            while not self._stop.is_set():
                try:
                    self._q.put(item, timeout=0.1)
                    break
                except queue.Full:
                    continue

    def next_batch(self):
        """Return the next (images, labels) batch, wrapping across epochs."""
        if self._prefetch == 0:
            return self._pin(self._raw_next())
        if self._thread is None:
            self._q = queue.Queue(maxsize=self._prefetch)
            self._thread = threading.Thread(target=self._stage_loop, daemon=True)
            self._thread.start()
        return self._q.get()

    def close(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
            self._thread = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
