"""Checkpoint layout (parity: SURVEY.md 5.4).

Model file: ``<train_dir> + "model_step_" + str(step)`` — plain string
concatenation exactly like ``sync_replicas_master_nn.py:331-336`` /
``distributed_worker.py:337-342`` (``torch.save(network.state_dict(), f)``), so
the polling evaluator's contract (``distributed_evaluator.py:76-88``) holds.
New: a sidecar ``..._optim`` with optimizer state, step, LR and RNG state for
true resume (the reference cannot resume).
"""
from __future__ import annotations

import os
from typing import Optional

import torch


def model_path(train_dir: str, step: int) -> str:
    return train_dir + "model_step_" + str(step)


def save_model(train_dir: str, step: int, network: torch.nn.Module) -> str:
    path = model_path(train_dir, step)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        torch.save(network.state_dict(), f)
    os.replace(tmp, path)  # atomic: the evaluator never sees a partial file
    return path


def save_sidecar(train_dir: str, step: int, optimizer=None, lr: Optional[float] = None, extra: Optional[dict] = None) -> str:
    path = model_path(train_dir, step) + "_optim"
    state = {"step": step, "lr": lr, "rng": torch.get_rng_state(),
             "optimizer": optimizer.state_dict() if optimizer is not None else None}
    if extra:
        state.update(extra)
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)
    return path


def latest_step(train_dir: str) -> Optional[int]:
    d = os.path.dirname(train_dir + "x") or "."
    prefix = os.path.basename(train_dir + "model_step_")
    best = None
    if not os.path.isdir(d):
        return None
    for name in os.listdir(d):
        if name.startswith(prefix) and not name.endswith(("_optim", ".tmp")):
            try:
                s = int(name[len(prefix):])
            except ValueError:
                continue
            best = s if best is None else max(best, s)
    return best


def load_model(train_dir: str, step: int, network: torch.nn.Module, map_location="cpu") -> None:
    with open(model_path(train_dir, step), "rb") as f:
        network.load_state_dict(torch.load(f, map_location=map_location))


def load_sidecar(train_dir: str, step: int, optimizer=None, map_location="cpu") -> Optional[dict]:
    path = model_path(train_dir, step) + "_optim"
    if not os.path.exists(path):
        return None
    state = torch.load(path, map_location=map_location, weights_only=False)
    if optimizer is not None and state.get("optimizer") is not None:
        optimizer.load_state_dict(state["optimizer"])
    return state
