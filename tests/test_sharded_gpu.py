"""Latitude-sharded forward (ONE forecast over N GPUs) against the single-GPU forward, for both transports of the halo
exchange: `peer` (kernels storing into the neighbours' memory over NVLink, CUDA IPC mappings, the whole step in ONE
CUDA graph) and `nccl` (NCCL send / recv between graph segments).  Every per-row computation is independent of how
rows are split over ranks, so the result must be IDENTICAL, bit for bit.

The multi-process cases need >= 2 (>= 8) CUDA devices (skipped otherwise; `gpurun --gpus 2` / `--gpus 8`); the
single-GPU case runs the peer kernels with the band as its own neighbour."""

import os
import socket

import pytest
import torch

from tests import fixtures as fx

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, h, w, one_gpu=False):
    import torch.distributed as dist

    import aurora_b200 as ab
    from aurora_b200 import sharding

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = 0 if one_gpu else rank
    torch.cuda.set_device(dev)
    if one_gpu:  # every rank is a process on the SAME GPU: gloo for the rendezvous, CUDA IPC between the processes
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = fx.CONFIGS["tiny_lora"]
    model = ab.Aurora(**fx.our_kwargs(cfg))
    model.load_state_dict(fx.make_state_dict(cfg, seed=31))
    model = model.to(f"cuda:{dev}").eval()
    # full 144-token windows at all three stages, zero padding in W at stage 3
    batch = fx.make_batch(cfg, h, w, levels=fx.LEVELS4, b=1, seed=31, rollout_step=1)
    ref = model.forward(batch) if rank == 0 else None
    ok, worst, notes = True, 0.0, []
    for mode in (("peer",) if one_gpu else ("nccl", "peer")):
        model.halo_mode = mode
        model.use_cuda_graph = False
        local = model.forward(batch, sharded=True)
        plans = local.slab_plans
        host = (lambda t: t.cpu()) if one_gpu else (lambda t: t)   # gloo gathers on the host
        full_surf = {k: sharding.gather_bands(host(v), plans, cfg.patch_size) for k, v in local.surf_vars.items()}
        full_atmos = {k: sharding.gather_bands(host(v), plans, cfg.patch_size) for k, v in local.atmos_vars.items()}
        torch.cuda.synchronize()
        keep = {k: v.clone() for k, v in local.atmos_vars.items()}
        # the same sharded step replayed from a CUDA graph: identical bits.  peer: ONE graph holds the whole step
        # including the 12 exchanges; nccl: 13 graph segments with the exchanges issued eagerly between them
        model.use_cuda_graph = True
        for _ in range(3):  # capture + replay, then replays only
            again = model.forward(batch, sharded=True)
            torch.cuda.synchronize()
            ok = ok and all(torch.equal(again.atmos_vars[k], keep[k]) for k in keep)
        entry = next(e for sig, e in model._engine._graphs.items() if sig[-1] == mode)
        n_graphs = sum(isinstance(i, torch.cuda.CUDAGraph) for i in entry["items"])
        ok = ok and n_graphs == (1 if mode == "peer" else 13)
        notes.append((mode, n_graphs))
        if rank == 0:
            worst_mode = 0.0
            for grp, got in ((ref.surf_vars, full_surf), (ref.atmos_vars, full_atmos)):
                for k, v in grp.items():
                    ok = ok and got[k].shape == v.shape
                    worst_mode = max(worst_mode, (got[k].to(v.device) - v).abs().max().item())
            notes.append((mode, "max |sharded - single|", worst_mode))
            worst = max(worst, worst_mode)
            ok = ok and worst == 0.0
        dist.barrier()
    q.put((rank, ok, worst, [p.rows for p in plans], notes))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,h,w", [(2, 192, 256), (8, 640, 256)])
def test_sharded_forward_equals_single_gpu(world, h, w):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, h, w)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=900) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    print(f"[sharded x{world}] slabs:", res[0][3][:2], "... max |sharded - single| =", res[0][2], res[0][4])


@pytest.mark.parametrize("world,h,w", [(2, 192, 256), (3, 240, 256), (8, 640, 256)])
def test_peer_halo_over_ipc_between_processes_on_one_gpu(world, h, w):
    """The real multi-process protocol on a SINGLE GPU: `world` processes share cuda:0 (the driver time-slices their
    contexts), rendezvous over gloo, map each other's halo buffers with CUDA IPC and run the sharded forward with the
    peer transport — push / release-flag / acquire-wait kernels between processes, one CUDA graph per step.  Must equal
    the unsharded forward bit for bit.  (NCCL refuses two ranks on one device, so only the peer transport runs here.)"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, h, w, True)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=900) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    print(f"[sharded x{world} on one GPU, IPC] max |sharded - single| =", res[0][2], res[0][4])


def test_peer_halo_kernels_on_one_gpu():
    """World size 1: the band is the whole grid and wraps onto itself, so the peer transport pushes into its own
    buffer — the same push / flag / wait kernels, no IPC.  Eager and single-graph replay must equal the plain
    (unsharded) forward bit for bit, on changing inputs."""
    import aurora_b200 as ab

    cfg = fx.CONFIGS["tiny_lora"]
    model = ab.Aurora(**fx.our_kwargs(cfg))
    model.load_state_dict(fx.make_state_dict(cfg, seed=17))
    model = model.to("cuda").eval()
    b1 = fx.make_batch(cfg, 192, 256, levels=fx.LEVELS4, b=1, seed=17, rollout_step=1)
    b2 = fx.make_batch(cfg, 192, 256, levels=fx.LEVELS4, b=1, seed=18, rollout_step=1)
    plain1 = {k: v.clone() for k, v in model.forward(b1).atmos_vars.items()}
    plain2 = {k: v.clone() for k, v in model.forward(b2).atmos_vars.items()}
    model.halo_mode = "peer"
    e1 = model.forward(b1, sharded=True).atmos_vars
    assert model._engine._peer is not None and model._engine._peer.world == 1
    for k in plain1:
        assert torch.equal(plain1[k], e1[k]), k
    model.use_cuda_graph = True
    g1 = model.forward(b1, sharded=True).atmos_vars
    g2 = model.forward(b2, sharded=True).atmos_vars
    g1b = model.forward(b1, sharded=True).atmos_vars
    entry = next(e for sig, e in model._engine._graphs.items() if sig[-1] == "peer")
    assert len(entry["items"]) == 1  # the whole step, exchanges included, is one graph
    for k in plain1:
        assert torch.equal(plain1[k], g1[k]) and torch.equal(plain2[k], g2[k]) and torch.equal(plain1[k], g1b[k]), k


def test_sharded_forward_from_pinned_host_memory_uploads_the_band():
    """A batch in PINNED host memory takes the overlapped upload in sharded mode too: only the rank's latitude band goes
    up, one DMA per (H, W) plane.  World size 1 (band = whole cropped grid, 193 -> 192 rows): same bits as the device path."""
    import aurora_b200 as ab
    from aurora_b200 import Batch

    cfg = fx.CONFIGS["tiny_lora"]
    model = ab.Aurora(**fx.our_kwargs(cfg))
    model.load_state_dict(fx.make_state_dict(cfg, seed=23))
    model = model.to("cuda").eval()
    model.halo_mode = "peer"
    batches = [fx.make_batch(cfg, 193, 256, levels=fx.LEVELS4, b=1, seed=50 + i, rollout_step=1) for i in range(3)]
    want = [{k: v.clone() for k, v in model.forward(b, sharded=True).atmos_vars.items()} for b in batches]
    pin = lambda d: {k: v.contiguous().pin_memory() for k, v in d.items()}  # noqa: E731
    pinned = [Batch(pin(b.surf_vars), pin(b.static_vars), pin(b.atmos_vars), b.metadata) for b in batches]
    torch.cuda.synchronize()
    preds = [model.forward(b, sharded=True) for b in pinned]   # no sync in between: uploads overlap the previous step
    torch.cuda.synchronize()
    assert model._engine._h2d is not None
    for p, w in zip(preds, want):
        assert next(iter(p.surf_vars.values())).shape[-2:] == (192, 256)
        for k in w:
            assert torch.equal(p.atmos_vars[k], w[k]), k
