"""Autoregressive roll-out (`aurora/rollout.py:14-49`): same generator semantics as the reference — the
consumer's code runs between steps — with the state kept on the model's device."""

from __future__ import annotations

import dataclasses
from typing import Generator

import torch

from aurora_b200.batch import Batch

__all__ = ["rollout"]


def rollout(model, batch: Batch, steps: int) -> Generator[Batch, None, None]:
    """Yield the prediction after each of `steps` model time steps, feeding predictions back as the
    newest history entry."""
    batch = model.batch_transform_hook(batch)
    p = next(model.parameters())
    batch = batch.type(p.dtype)
    batch = batch.crop(model.patch_size)
    batch = batch.to(p.device)
    for _ in range(steps):
        pred = model.forward(batch)
        yield pred
        batch = dataclasses.replace(
            pred,
            surf_vars={k: torch.cat([batch.surf_vars[k][:, 1:], v], dim=1) for k, v in pred.surf_vars.items()},
            atmos_vars={k: torch.cat([batch.atmos_vars[k][:, 1:], v], dim=1) for k, v in pred.atmos_vars.items()},
        )
