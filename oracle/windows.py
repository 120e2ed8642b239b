"""ORACLE (test infrastructure, never imported by the product): integer index arithmetic of the 3-D
shifted-window attention, restated in closed form with numpy.

Restates, as one gather map and one group-id map, the chain
`torch.roll(-ss)` -> `pad_3d` -> `window_partition_3d` (and its inverse) and the mask builder of the
reference:

* `maybe_adjust_windows`            aurora/model/util.py:53-71
* `get_three_sidded_padding`        aurora/model/swin3d.py:177-194, 250-269
* `window_partition_3d`             aurora/model/swin3d.py:197-214
* roll / pad / partition in a block aurora/model/swin3d.py:470-486, reverse :491-503
* `compute_3d_shifted_window_mask`  aurora/model/swin3d.py:303-360 (+ merge groups :288-300)

Pinned against the reference itself (index tensors pushed through the reference functions) by
tests/golden/make_golden.py -> tests/golden/windows_*.npz; see tests/test_oracle_windows.py.
"""

from __future__ import annotations

import functools

import numpy as np

PAD_GROUP = 27  # group id the reference assigns to zero-padded tokens (swin3d.py:348-352)


def adjust_windows(ws0, ss0, res):
    """Clamp the window to the resolution and drop the shift on clamped axes (util.py:53-71)."""
    ws, ss = list(ws0), list(ss0)
    for a in range(3):
        if res[a] <= ws0[a]:
            ws[a] = res[a]
            ss[a] = 0
    return tuple(ws), tuple(ss)


def pad_lo_hi(res, ws):
    """Per-axis (front, back) zero padding up to a multiple of the window (swin3d.py:481, 177-194)."""
    out = []
    for a in range(3):
        pad = (-res[a]) % ws[a]
        lo = pad // 2
        out.append((lo, pad - lo))
    return out


def window_gather_map(res, ws0, ss0):
    """For every (window, in-window token) the flat source token `(c*H + h)*W + w` it is read from and
    written back to, or -1 for a zero-padded position.

    Returns (idx[nW, N] int64, ws, ss, n_windows_per_axis).
    """
    ws, ss = adjust_windows(ws0, ss0, res)
    pads = pad_lo_hi(res, ws)
    n = [(res[a] + pads[a][0] + pads[a][1]) // ws[a] for a in range(3)]
    # Coordinates in the padded, shifted frame for every (window index, in-window index) pair.
    coords = []
    valid = None
    for a in range(3):
        k = np.arange(n[a])[:, None]
        i = np.arange(ws[a])[None, :]
        q = k * ws[a] + i - pads[a][0]  # shifted-unpadded coordinate, may fall outside [0, res)
        ok = (q >= 0) & (q < res[a])
        src = (q + ss[a]) % res[a]  # roll(x, -ss)[q] == x[(q + ss) % n]
        coords.append((src, ok))
    (sc, okc), (sh, okh), (sw, okw) = coords
    # window id = (c1*n1 + h1)*n2 + w1 ; token id = (ic*ws1 + ih)*ws2 + iw
    src = (sc[:, None, None, :, None, None] * res[1] + sh[None, :, None, None, :, None]) * res[2] + sw[
        None, None, :, None, None, :
    ]
    valid = okc[:, None, None, :, None, None] & okh[None, :, None, None, :, None] & okw[None, None, :, None, None, :]
    idx = np.where(valid, src, -1).reshape(n[0] * n[1] * n[2], ws[0] * ws[1] * ws[2]).astype(np.int64)
    return idx, ws, ss, tuple(n)


def _axis_bucket(q, res_a, ws_a, ss_a):
    """Slice index (0, 1, 2) of a shifted-frame coordinate along one axis (swin3d.py:333-342).  With
    ss == 0 the first two Python slices `[0:-ws]`... degenerate so that everything lands in the last."""
    if ss_a == 0:
        # slice(-0, None) == slice(0, None) covers the whole axis and is assigned last.
        return np.full_like(q, 2)
    b = np.where(q < res_a - ws_a, 0, np.where(q < res_a - ss_a, 1, 2))
    return b


def window_group_ids(res, ws0, ss0, warped=True):
    """Group id of every (window, token): tokens attend to each other iff ids are equal.  None when the
    block is not shifted (the reference then applies NO mask, even with padding: swin3d.py:476-478)."""
    ws, ss = adjust_windows(ws0, ss0, res)
    if all(s == 0 for s in ss):
        return None
    pads = pad_lo_hi(res, ws)
    n = [(res[a] + pads[a][0] + pads[a][1]) // ws[a] for a in range(3)]
    bs, oks = [], []
    for a in range(3):
        k = np.arange(n[a])[:, None]
        i = np.arange(ws[a])[None, :]
        q = k * ws[a] + i - pads[a][0]
        oks.append((q >= 0) & (q < res[a]))
        b = _axis_bucket(np.clip(q, 0, res[a] - 1), res[a], ws[a], ss[a])
        if a == 2 and warped:
            b = np.where(b == 1, 2, b)  # left/right edges are connected (swin3d.py:288-300, 344-346)
        bs.append(b)
    g = (
        9 * bs[0][:, None, None, :, None, None]
        + 3 * bs[1][None, :, None, None, :, None]
        + bs[2][None, None, :, None, None, :]
    )
    valid = (
        oks[0][:, None, None, :, None, None] & oks[1][None, :, None, None, :, None] & oks[2][None, None, :, None, None, :]
    )
    g = np.where(valid, g, PAD_GROUP)
    return g.reshape(n[0] * n[1] * n[2], ws[0] * ws[1] * ws[2]).astype(np.uint8)


def shifted_window_mask(res, ws0, ss0, warped=True, dtype=np.float32):
    """(nW, N, N) additive mask: 0 inside a group, -100 across groups (swin3d.py:357-358).  Memoised per
    geometry like the reference's `lru_cache` on `compute_3d_shifted_window_mask` (swin3d.py:303); the
    returned array is read-only."""
    return _shifted_window_mask(tuple(res), tuple(ws0), tuple(ss0), bool(warped), np.dtype(dtype))


@functools.lru_cache(maxsize=16)
def _shifted_window_mask(res, ws0, ss0, warped, dtype):
    g = window_group_ids(res, ws0, ss0, warped)
    if g is None:
        return None
    same = g[:, :, None] == g[:, None, :]
    m = np.where(same, dtype.type(0.0), dtype.type(-100.0))
    m.setflags(write=False)
    return m
