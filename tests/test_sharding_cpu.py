"""Latitude sharding host logic on CPU: slab planner and the halo exchange over two / three gloo ranks."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aurora_b200 import sharding as S


def test_plan_slabs_production_grid():
    plans = S.plan_slabs(180, 3, 8)
    assert [p.rows[0][1] for p in plans] == [24, 24, 24, 24, 24, 20, 20, 20]
    assert sum(p.rows[0][1] for p in plans) == 180
    for p in plans:
        for s in range(3):
            start, cnt = p.rows[s]
            assert start == p.rows[0][0] // 2**s and cnt == p.rows[0][1] // 2**s and cnt >= S.HALO
        assert p.global_h == (180, 90, 45)
    # contiguous cover at every stage
    for s in range(3):
        pos = 0
        for p in plans:
            assert p.rows[s][0] == pos
            pos += p.rows[s][1]
        assert pos == 180 // 2**s
    assert plans[3].image_rows(4) == (72 * 4, 24 * 4)


def test_plan_slabs_rejects_bad_splits():
    with pytest.raises(NotImplementedError):
        S.plan_slabs(150, 3, 2)   # 150 -> 75 -> 38: odd merge inside the U-Net
    with pytest.raises(ValueError):
        S.plan_slabs(180, 3, 16)  # bands too thin for the halo


def test_exchange_halo_single_process_is_cyclic():
    x = torch.arange(2 * 7 * 3, dtype=torch.float32).view(2, 7, 3)
    h = S.exchange_halo(x, 2)
    assert torch.equal(h[0], x[:, 5:]) and torch.equal(h[1], x[:, :2])


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    c, h, x, halo = 3, 8 * world, 5, 3
    full = torch.randn(c, h, x)
    rows = h // world
    local = full[:, rank * rows:(rank + 1) * rows].contiguous()
    got = S.exchange_halo(local, halo)
    top = [(rank * rows - halo + i) % h for i in range(halo)]
    bot = [((rank + 1) * rows + i) % h for i in range(halo)]
    ok = torch.equal(got[0], full[:, top]) and torch.equal(got[1], full[:, bot])
    # the K | V form the engine uses: a [C, rows, W, 3D] projection, of which only columns [D, 3D) are exchanged
    w_tok, d = 4, 2
    full4 = torch.randn(c, h, w_tok, 3 * d)
    got4 = S.exchange_halo(full4[:, rank * rows:(rank + 1) * rows].contiguous(), halo, col_from=d)
    ok = ok and got4.shape == (2, c, halo, w_tok, 2 * d)
    ok = ok and torch.equal(got4[0], full4[:, top][..., d:]) and torch.equal(got4[1], full4[:, bot][..., d:])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_halo_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


@pytest.mark.parametrize("h,world", [(180, 8), (90, 8), (45, 8), (48, 2), (24, 2), (160, 8), (40, 8), (78, 3)])
@pytest.mark.parametrize("shifted", [False, True])
def test_halo_needs_cover_exactly_the_rows_the_windows_reach(h, world, shifted):
    """`halo_needs` against the oracle's window gather map (oracle/windows.py, bit-exact vs the reference): for every
    band, the rows of all windows that contain at least one owned row, minus the owned rows, must be exactly the
    `above` rows just above plus the `below` rows just below the band (nearest-first, no gaps needed beyond them)."""
    import numpy as np

    from oracle import windows as W

    ws0, c, w = (2, 6, 12), 2, 12
    ss0 = (1, 3, 6) if shifted else (0, 0, 0)
    idx, ws, ss, _ = W.window_gather_map((c, h, w), ws0, ss0)
    rows_of_window = [set(((t // w) % h) for t in win if t >= 0) for win in np.asarray(idx)]
    base, rem = divmod(h, world)
    start = 0
    for r in range(world):
        cnt = base + (1 if r < rem else 0)
        owned = set(range(start, start + cnt))
        needed = set()
        for rows in rows_of_window:
            if rows & owned:
                needed |= rows - owned
        above, below = S.halo_needs(h, ws0[1], ss0[1], start, cnt)
        assert 0 <= above < ws0[1] and 0 <= below < ws0[1]
        got = {(start - 1 - i) % h for i in range(above)} | {(start + cnt + i) % h for i in range(below)}
        assert needed <= got, (r, sorted(needed), sorted(got))
        # tight: the farthest row on each side is really needed
        if above:
            assert (start - above) % h in needed
        if below:
            assert (start + cnt + below - 1) % h in needed
        start += cnt
