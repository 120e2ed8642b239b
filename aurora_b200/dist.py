"""Multi-GPU plumbing shared by both parallel modes: one process per GPU, `torch.distributed` for rendezvous, a barrier
and a max-reduction of device timings.  Independent forecasts (replicas) need no data-path collective; the latitude
sharding of ONE forecast — the default of `bench.py --gpus N` — and its peer-memory / NCCL halo exchange live in
`aurora_b200/sharding.py` and `csrc/halo.cu`."""

from __future__ import annotations

import os
from typing import Sequence

import torch
import torch.distributed as dist

__all__ = ["env_world", "init_process_group", "max_over_ranks", "shard_indices"]


def env_world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str, device: torch.device | None = None) -> bool:
    """Initialise the default process group when launched under torchrun; returns whether distributed."""
    rank, world, _ = env_world()
    if world <= 1:
        return False
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    kw = {}
    if backend == "nccl" and device is not None:
        kw["device_id"] = device
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return True


def max_over_ranks(values: Sequence[float], device: torch.device | str = "cpu") -> list[float]:
    """Element-wise maximum of per-rank measurements (timings are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def shard_indices(n_units: int, rank: int, world: int) -> range:
    """Contiguous, balanced partition of independent work units (forecasts) over ranks."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
