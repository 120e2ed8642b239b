"""Autoregressive roll-out with the reference's generator contract (`aurora/rollout.py:14-49`): every
step's prediction is yielded — the caller's code runs between steps — and then becomes the newest
entry of the history window.  The state stays on the model's device for the whole roll-out."""

from __future__ import annotations

from typing import Generator, Mapping

import torch

from aurora_b200.batch import Batch

__all__ = ["rollout"]


def _push(history: Mapping[str, torch.Tensor], newest: Mapping[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """Drop the oldest time slice (dim 1) of every variable and append the prediction."""
    return {name: torch.cat((history[name][:, 1:], step), dim=1) for name, step in newest.items()}


def _initial_state(model, batch: Batch) -> Batch:
    first_param = next(model.parameters())
    state = model.batch_transform_hook(batch)  # may add / remove variables (wave model)
    return state.type(first_param.dtype).crop(model.patch_size).to(first_param.device)


def rollout(model, batch: Batch, steps: int) -> Generator[Batch, None, None]:
    """Yield `steps` successive predictions of `model`, starting from `batch`.

    Static variables and metadata (time advanced by the model time step, `rollout_step` incremented)
    are taken from each prediction, exactly as the reference does.
    """
    state = _initial_state(model, batch)
    remaining = int(steps)
    while remaining > 0:
        prediction = model.forward(state)
        yield prediction
        state = Batch(
            surf_vars=_push(state.surf_vars, prediction.surf_vars),
            static_vars=prediction.static_vars,
            atmos_vars=_push(state.atmos_vars, prediction.atmos_vars),
            metadata=prediction.metadata,
        )
        remaining -= 1
