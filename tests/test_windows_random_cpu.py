"""Window / shift / mask index arithmetic on RANDOM geometries (no GPU): the C library's `__host__ __device__`
routine (the one the attention kernels use) against the oracle's closed form, structural invariants, and — when the
reference checkout is present (build container) — the reference's own roll -> pad -> window_partition_3d and
compute_3d_shifted_window_mask run live.  Extends the 34 stored golden geometries to a few hundred."""

import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from aurora_b200 import cabi
from oracle import windows as W

WS0 = (2, 6, 12)
REF = Path("/root/reference")


def _geometries(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for _ in range(n):
        res = (int(rng.integers(1, 7)), int(rng.integers(1, 41)), int(rng.integers(1, 61)))
        out.append((res, bool(rng.integers(0, 2)), bool(rng.integers(0, 2))))
    return out


@pytest.mark.parametrize("chunk", range(6))
def test_host_map_equals_oracle_and_is_a_permutation(chunk):
    for res, shifted, warped in _geometries(40, 100 + chunk):
        ss0 = tuple(s // 2 for s in WS0) if shifted else (0, 0, 0)
        idx_o, ws, ss, nwin = W.window_gather_map(res, WS0, ss0)
        grp_o = W.window_group_ids(res, WS0, ss0, warped)
        idx_c, grp_c = cabi.window_index_map_host(res, WS0, ss0, warped)
        assert idx_c.shape == idx_o.shape, (res, shifted)
        np.testing.assert_array_equal(idx_c, idx_o.astype(np.int32), err_msg=str((res, shifted, warped)))
        if grp_o is not None:
            np.testing.assert_array_equal(grp_c, grp_o, err_msg=str((res, shifted, warped)))
        # every real token exactly once, the rest is padding
        real = idx_c[idx_c >= 0]
        n_tok = res[0] * res[1] * res[2]
        assert real.size == n_tok and np.array_equal(np.sort(real), np.arange(n_tok)), (res, shifted)
        assert (idx_c < 0).sum() == idx_c.size - n_tok
        if grp_o is not None:
            assert (grp_c[idx_c < 0] == W.PAD_GROUP).all() and (grp_c[idx_c >= 0] < W.PAD_GROUP).all()
        assert cabi.window_geometry(res, WS0, ss0)[:2] == idx_o.shape


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("chunk", range(4))
def test_host_map_equals_reference_live(chunk):
    sys.path[:0] = [str(REF), str(Path(__file__).parent / "_shims")]
    try:
        from aurora.model import swin3d as ref
        from aurora.model.util import maybe_adjust_windows
    finally:
        del sys.path[:2]
    compared = 0
    for res, shifted, warped in _geometries(30, 200 + chunk):
        c, h, w = res
        ss0 = tuple(s // 2 for s in WS0) if shifted else (0, 0, 0)
        ws, ss = maybe_adjust_windows(WS0, ss0, res)
        x = (torch.arange(c * h * w, dtype=torch.float64) + 1).view(1, c, h, w, 1)
        if any(ss):
            x = torch.roll(x, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        x = ref.pad_3d(x, ((-c) % ws[0], (-h) % ws[1], (-w) % ws[2]))
        idx_r = ref.window_partition_3d(x, ws).reshape(-1, ws[0] * ws[1] * ws[2]).long().numpy() - 1
        idx_c, grp_c = cabi.window_index_map_host(res, WS0, ss0, warped)
        np.testing.assert_array_equal(idx_c, idx_r.astype(np.int32), err_msg=str((res, shifted)))
        if any(ss):
            ref.compute_3d_shifted_window_mask.cache_clear()
            try:
                mask, img = ref.compute_3d_shifted_window_mask(c, h, w, ws, ss, torch.device("cpu"), torch.float32, warped)
            except RuntimeError:  # the reference's own `.view` fails on some degenerate grids (swin3d.py:355)
                continue
            grp_r = ref.window_partition_3d(img, ws).reshape(-1, ws[0] * ws[1] * ws[2]).to(torch.uint8).numpy()
            np.testing.assert_array_equal(grp_c, grp_r, err_msg=str((res, shifted, warped)))
            same = torch.from_numpy(grp_c)[:, :, None] == torch.from_numpy(grp_c)[:, None, :]
            assert torch.equal(mask, torch.where(same, 0.0, -100.0)), (res, warped)
        compared += 1
    assert compared >= 20
