"""Latitude sharding of ONE forecast over the GPUs of a node.

Every rank owns a contiguous band of latitude token rows at every U-Net stage.  Row-wise work (all GEMMs,
adaLN, MLP, Perceiver encoder / decoder, patch merge / split, patch I/O) touches only local rows.  Window
attention is the only exchange step: before each attention call the ranks swap `HALO` rows of the qkv
projection with both neighbours (cyclic in latitude, because the shifted windows wrap), every rank then
computes all windows touching its band (boundary windows are computed on both sides) and writes only its own
rows — one send + one receive per neighbour per block, no all-reduce anywhere.

Two transports for that exchange: `PeerHalo` — kernels of libaurora_b200.so that store the rows straight into the
neighbours' memory over NVLink (CUDA IPC mappings) and hand over with release / acquire flags, all inside the
step's CUDA graph — and `exchange_halo`, NCCL send / recv issued from the host between graph segments.
"""

from __future__ import annotations

import dataclasses
from typing import Optional

import torch
import torch.distributed as dist

__all__ = ["HALO", "SlabPlan", "plan_slabs", "halo_needs", "exchange_halo", "gather_bands", "PeerHalo"]

HALO = 5  # window height 6: a window reaches at most 5 rows into a neighbouring band


@dataclasses.dataclass(frozen=True)
class SlabPlan:
    """Row ranges of one rank.  `rows[s] = (start, count)` in token rows of U-Net stage `s`."""

    rank: int
    world: int
    rows: tuple[tuple[int, int], ...]
    global_h: tuple[int, ...]  # global token rows per stage

    def image_rows(self, patch: int) -> tuple[int, int]:
        s, c = self.rows[0]
        return s * patch, c * patch


def plan_slabs(h0: int, n_stages: int, world: int) -> list[SlabPlan]:
    """Split `h0` stage-0 token rows into `world` bands whose sizes are multiples of 2^(n_stages-1), so that
    every 2x2 patch merge / split stays inside a band; bands differ by at most one unit."""
    unit = 2 ** (n_stages - 1)
    if h0 % unit != 0:
        raise NotImplementedError(
            f"latitude sharding needs the token grid height ({h0}) to be a multiple of {unit} "
            f"(no odd-size patch merging inside the U-Net)")
    n_units = h0 // unit
    if n_units < world:
        raise ValueError(f"cannot split {h0} token rows over {world} ranks")
    base, rem = divmod(n_units, world)
    if base < HALO:
        raise ValueError(
            f"bands would own {base} rows at the deepest stage, fewer than the {HALO}-row halo their neighbours need; "
            f"use at most {n_units // HALO} ranks")
    plans, start = [], 0
    for r in range(world):
        cnt = (base + (1 if r < rem else 0)) * unit
        rows = tuple((start // 2**s, cnt // 2**s) for s in range(n_stages))
        plans.append(SlabPlan(r, world, rows, tuple(h0 // 2**s for s in range(n_stages))))
        start += cnt
    return plans


def halo_needs(h: int, window_h: int, shift_h: int, h_begin: int, h_rows: int) -> tuple[int, int]:
    """How many token rows ABOVE and BELOW the band `[h_begin, h_begin + h_rows)` (cyclic in the global height `h`) the
    attention windows touching the band reach: `(rows_above, rows_below)`, each in `0 .. window_h - 1`.  Same geometry
    as `csrc/window_index.cuh`: two-sided zero padding to a multiple of the window (front = pad // 2), cyclic shift of
    the token grid, windows never wrap across the padded frame (swin3d.py:470-503).  A band that covers the whole grid
    needs nothing."""
    if h <= window_h:  # maybe_adjust_windows: one window over the whole axis, no shift
        window_h, shift_h = h, 0
    pad = (-h) % window_h
    lo = pad // 2
    above = below = 0
    owned = lambda r: (r - h_begin) % h < h_rows  # noqa: E731
    for kh in range((h + pad) // window_h):
        rows = [(q + shift_h) % h for q in range(kh * window_h - lo, (kh + 1) * window_h - lo) if 0 <= q < h]
        if not any(owned(r) for r in rows):
            continue
        for r in rows:
            if owned(r):
                continue
            up = (h_begin - r) % h                 # 1 = the row just above the band
            down = (r - (h_begin + h_rows)) % h + 1  # 1 = the row just below
            if up <= down:
                above = max(above, up)
            else:
                below = max(below, down)
    return above, below


def exchange_halo(local: torch.Tensor, halo: int, out: Optional[torch.Tensor] = None, group=None,
                  col_from: int = 0) -> torch.Tensor:
    """`local` is one rank's band `[C, rows, X]` (X = W * channels, contiguous) or `[C, rows, W, K]`, of which only the
    columns `[col_from, K)` of every token are exchanged (k | v of a qkv projection).  Returns `[2, C, halo, X]`
    (`[2, C, halo, W, K - col_from]`): index 0 = the `halo` rows just above the band (the previous rank's last rows),
    index 1 = the rows just below (the next rank's first rows); bands are cyclic neighbours.  One send + one receive
    per neighbour."""
    c, rows = local.shape[:2]
    assert rows >= halo, f"band of {rows} rows cannot serve a {halo}-row halo"
    if local.dim() == 4:
        local = local[..., col_from:]
    else:
        assert col_from == 0
    if out is None:
        out = torch.empty(2, c, halo, *local.shape[2:], dtype=local.dtype, device=local.device)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    first = local[:, :halo].contiguous()
    last = local[:, rows - halo:].contiguous()
    if world == 1:
        out[0].copy_(last)
        out[1].copy_(first)
        return out
    rank = dist.get_rank(group)
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    # Posting order matters when prev == nxt (two ranks): the peer posts (first -> me, last -> me), so receive in
    # the same order: its first rows are my bottom halo, its last rows my top halo.
    ops = [
        dist.P2POp(dist.isend, first, prev, group),
        dist.P2POp(dist.isend, last, nxt, group),
        dist.P2POp(dist.irecv, out[1], nxt, group),
        dist.P2POp(dist.irecv, out[0], prev, group),
    ]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return out


def gather_bands(local: torch.Tensor, plans: list[SlabPlan], patch: int, dim: int = -2, group=None) -> torch.Tensor:
    """All-gather the latitude bands of an output field (dimension `dim` = latitude) into the full field."""
    world = len(plans)
    if world == 1:
        return local
    sizes = [p.image_rows(patch)[1] for p in plans]
    dim = dim % local.dim()
    pieces = []
    for p, n in zip(plans, sizes):
        shape = list(local.shape)
        shape[dim] = n
        pieces.append(torch.empty(shape, dtype=local.dtype, device=local.device))
    dist.all_gather(pieces, local.contiguous(), group=group)
    return torch.cat(pieces, dim=dim)


class PeerHalo:
    """Peer-memory transport of the halo exchange (protocol: `aurora_b200/csrc/halo.cu`).

    One device buffer per rank — control words + two parity regions, each holding the rows above (side 0) and below
    (side 1) the band — mapped by both neighbours through CUDA IPC.  `exchange` launches two kernels on the current
    stream (push into the neighbours, wait for the neighbours' pushes) and returns this rank's halo `[2, C, halo, X]`;
    nothing synchronises the host, so the calls can be captured in a CUDA graph.  Collective: every rank of `group`
    must construct it, and call `exchange` the same number of times, in the same order.  A world of one rank is its own
    neighbour (the band wraps onto itself), which exercises the same kernels without IPC.
    """

    def __init__(self, device: torch.device, side_bytes_max: int, group=None) -> None:
        from aurora_b200 import cabi

        self._cabi = cabi
        self.group = group
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        self.world, self.rank = world, rank
        self.ctrl_bytes = cabi.AB_HALO_CTRL_BYTES
        self.region_bytes = (2 * side_bytes_max + 255) // 256 * 256
        self.buf = torch.zeros(self.ctrl_bytes + 2 * self.region_bytes, dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)  # zero flags are in memory before any peer can write a round number
        self.base = self.buf.data_ptr()
        self._opened: dict[int, int] = {}
        if world == 1:
            self.above = self.below = self.base
        else:
            handle, offset = cabi.ipc_export(self.buf)
            table: list = [None] * world
            dist.all_gather_object(table, (handle, offset), group=group)

            def peer_address(r: int) -> int:
                if r == rank:
                    return self.base
                if r not in self._opened:
                    self._opened[r] = cabi.ipc_open(table[r][0])
                return self._opened[r] + table[r][1]

            self.above = peer_address((rank - 1) % world)
            self.below = peer_address((rank + 1) % world)
            dist.barrier(group=group)  # every buffer is zeroed and mapped before the first push of any rank
        self.index = 0

    def begin_step(self) -> None:
        """Parity restarts with every model step (a step has an even number of exchanges: 2 x sum(encoder depths))."""
        assert self.index % 2 == 0, "odd number of halo exchanges in the previous step"
        self.index = 0

    def exchange(self, local: torch.Tensor, halo: int, rows_to_above: int, rows_to_below: int,
                 col_from: int = 0, wait: bool = True) -> torch.Tensor:
        """`local` = this rank's band `[C, rows, W, K]` (contiguous, 2-byte elements).  Sends columns `[col_from, K)`
        of its first `rows_to_above` rows to the rank above and of its last `rows_to_below` rows to the rank below
        (what THEIR windows reach into this band, `halo_needs`), and returns this rank's halo
        `[2, C, halo, W, K - col_from]`: index 0 = rows above the band (the nearest ones valid, at the END of the
        `halo` rows), index 1 = rows below (nearest first); valid once the stream reaches this point (`wait=True`) or
        once the consumer has waited on the control words itself (`wait=False`)."""
        cabi = self._cabi
        c, rows, w, k = local.shape
        kk = k - col_from
        side = c * halo * w * kk * local.element_size()
        if 2 * side > self.region_bytes or side % 16 != 0:
            raise ValueError(f"halo of {side} bytes per side does not fit the {self.region_bytes}-byte region")
        parity = self.index & 1
        self.index += 1
        region = self.ctrl_bytes + parity * self.region_bytes
        cabi.halo_push(local, above_slot=self.above + region + side, below_slot=self.below + region,
                       above_flag=self.above + 4, below_flag=self.below, ctrl=self.base, slot_rows=halo,
                       rows_to_above=rows_to_above, rows_to_below=rows_to_below, col_from=col_from)
        if wait:  # `wait=False`: the consumer (the attention kernel, given `self.base` as halo_ctrl) waits itself
            cabi.halo_wait(self.base)
        return self.buf[region:region + 2 * side].view(local.dtype).view(2, c, halo, w, kk)

    def descriptor(self, shape: tuple, dtype: torch.dtype, halo: int, rows_to_above: int, rows_to_below: int,
                   col_from: int = 0):
        """The same exchange as a launch descriptor for `ab_swin_block` (which fills in the source pointer and issues
        push + wait between the qkv projection and the attention): returns `(AbHaloPush, halo view)` and advances the
        parity exactly like `exchange`."""
        cabi = self._cabi
        c, rows, w, k = shape
        es = torch.empty((), dtype=dtype).element_size()
        kk = k - col_from
        side = c * halo * w * kk * es
        if 2 * side > self.region_bytes or side % 16 != 0:
            raise ValueError(f"halo of {side} bytes per side does not fit the {self.region_bytes}-byte region")
        parity = self.index & 1
        self.index += 1
        region = self.ctrl_bytes + parity * self.region_bytes
        a = cabi.AbHaloPush()
        a.above_slot, a.below_slot = self.above + region + side, self.below + region
        a.above_flag, a.below_flag, a.ctrl = self.above + 4, self.below, self.base
        a.c, a.rows, a.w, a.slot_rows = c, rows, w, halo
        a.rows_to_above, a.rows_to_below = rows_to_above, rows_to_below
        a.src_tok_bytes, a.tok_off_bytes, a.tok_bytes = k * es, col_from * es, kk * es
        return a, self.buf[region:region + 2 * side].view(dtype).view(2, c, halo, w, kk)

    def close(self) -> None:
        for base in self._opened.values():
            self._cabi.ipc_close(base)
        self._opened.clear()
