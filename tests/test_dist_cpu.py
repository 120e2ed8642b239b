"""N > 1 host logic on CPU: two gloo ranks, rendezvous on 127.0.0.1."""

import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from aurora_b200 import dist as abd


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q) -> None:
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    assert abd.init_process_group("gloo")
    vals = abd.max_over_ranks([10.0 + rank, 5.0 - rank])
    dist.barrier()
    units = list(abd.shard_indices(7, rank, world))
    q.put((rank, vals, units))
    dist.destroy_process_group()


def test_two_rank_gloo_reduction_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == out[1][1] == [11.0, 5.0]
    assert out[0][2] + out[1][2] == list(range(7))


def test_single_process_defaults():
    assert abd.env_world()[1] >= 1
    assert abd.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]
    assert list(abd.shard_indices(5, 0, 1)) == [0, 1, 2, 3, 4]
