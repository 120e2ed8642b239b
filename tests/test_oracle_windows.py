"""The oracle's closed-form window gather map / mask group ids are bit-exact against the reference's
roll -> pad -> window_partition_3d and compute_3d_shifted_window_mask (golden: tests/golden/windows.npz,
produced from the imported reference by tests/golden/make_golden.py)."""

import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import windows as W

GOLD = Path(__file__).parent / "golden"
WS0 = (2, 6, 12)


def _cases():
    z = np.load(GOLD / "windows.npz")
    for key in z.files:
        if key.startswith("idx_"):
            yield key[4:]


@pytest.mark.parametrize("tag", sorted(_cases()))
def test_gather_map_and_groups_bit_exact(tag):
    z = np.load(GOLD / "windows.npz")
    dims, sh, wp = tag.split("_")
    res = tuple(int(v) for v in dims.split("x"))
    shifted, warped = sh == "s", wp == "w"
    ss0 = tuple(s // 2 for s in WS0) if shifted else (0, 0, 0)
    idx, ws, ss, n = W.window_gather_map(res, WS0, ss0)
    assert idx.dtype == np.int64
    np.testing.assert_array_equal(idx, z[f"idx_{tag}"].astype(np.int64))
    g = W.window_group_ids(res, WS0, ss0, warped)
    if f"grp_{tag}" in z.files:
        np.testing.assert_array_equal(g, z[f"grp_{tag}"])
    else:
        assert g is None
    # every real token appears exactly once
    flat = idx[idx >= 0]
    assert flat.size == res[0] * res[1] * res[2] and np.unique(flat).size == flat.size


def test_production_shapes_by_checksum():
    hashes = json.loads((GOLD / "windows_hashes.json").read_text())
    for key, want in hashes.items():
        kind, dims, sh, _ = key.split("_")
        res = tuple(int(v) for v in dims.split("x"))
        ss0 = tuple(s // 2 for s in WS0) if sh == "s" else (0, 0, 0)
        if kind == "idx":
            got = W.window_gather_map(res, WS0, ss0)[0].astype(np.int32)
        else:
            got = W.window_group_ids(res, WS0, ss0, True)
        assert hashlib.sha256(got.tobytes()).hexdigest() == want, key


def test_mask_values():
    m = W.shifted_window_mask((4, 12, 24), WS0, (1, 3, 6))
    assert set(np.unique(m).tolist()) == {-100.0, 0.0}
    assert W.shifted_window_mask((4, 12, 24), WS0, (0, 0, 0)) is None
