"""`Batch` / `Metadata` host logic (no GPU): validation and error behaviour of `aurora/batch.py:24-190`, the
normalisation constants of `aurora/normalisation.py`, and — when the reference checkout is present (build
container only) — a live cross-check of normalise / unnormalise / crop against the reference classes."""

import sys
from datetime import datetime
from pathlib import Path

import pytest
import torch

from aurora_b200 import Batch, Metadata
from aurora_b200 import stats
from tests import fixtures as fx

REF = Path("/root/reference")


def _meta(h=17, w=32, **kw):
    base = dict(lat=torch.linspace(90, -90, h), lon=torch.linspace(0, 360, w + 1)[:-1],
                time=(datetime(2020, 6, 1, 12, 0),), atmos_levels=(100, 250, 500, 850))
    base.update(kw)
    return Metadata(**base)


@pytest.mark.parametrize("kw,msg", [
    (dict(lat=torch.linspace(91, -90, 17)), "Latitudes must be in the range"),
    (dict(lon=torch.linspace(0, 360, 32)), "Longitudes must be in the range"),
    (dict(lat=torch.linspace(-90, 90, 17)), "strictly decreasing"),
    (dict(lon=torch.linspace(0, 360, 33)[:-1].flip(0)), "strictly increasing"),
    (dict(lat=torch.linspace(90, -90, 17)[:, None].expand(17, 32)), "both be vectors or both be matrices"),
])
def test_metadata_validation_errors(kw, msg):
    with pytest.raises(ValueError, match=msg):
        _meta(**kw)


def test_crop_drops_one_latitude_row_and_rejects_more():
    cfg = fx.CONFIGS["tiny"]
    b = fx.make_batch(cfg, 17, 32, levels=fx.LEVELS4)
    c = b.crop(4)
    assert c.spatial_shape == (16, 32) and c.metadata.lat.shape == (16,)
    assert torch.equal(c.surf_vars["2t"], b.surf_vars["2t"][..., :-1, :])
    assert torch.equal(c.atmos_vars["t"], b.atmos_vars["t"][..., :-1, :])
    assert torch.equal(c.static_vars["z"], b.static_vars["z"][:-1])
    assert c.crop(4) is c  # already a multiple: unchanged
    with pytest.raises(ValueError, match="at most be one latitude too many"):
        fx.make_batch(cfg, 18, 32, levels=fx.LEVELS4).crop(4)
    with pytest.raises(ValueError, match="Width"):
        fx.make_batch(cfg, 16, 30, levels=fx.LEVELS4).crop(4)


def test_normalise_roundtrip_and_overrides():
    cfg = fx.CONFIGS["tiny"]
    b = fx.make_batch(cfg, 16, 32, levels=fx.LEVELS4)
    n = b.normalise()
    loc, sc = stats.surf_stats_of("2t")
    assert torch.allclose(n.surf_vars["2t"], (b.surf_vars["2t"] - loc) / sc)
    locs, scs = stats.atmos_stats_of("q", fx.LEVELS4)
    want = (b.atmos_vars["q"] - torch.tensor(locs)[:, None, None]) / torch.tensor(scs)[:, None, None]
    assert torch.allclose(n.atmos_vars["q"], want)
    back = n.unnormalise()
    for k in b.surf_vars:
        assert torch.allclose(back.surf_vars[k], b.surf_vars[k], rtol=1e-5, atol=1e-5 * abs(stats.surf_stats_of(k)[0]))
    over = {"2t": (1.0, 2.0)}
    assert torch.allclose(b.normalise(surf_stats=over).surf_vars["2t"], (b.surf_vars["2t"] - 1.0) / 2.0)
    d = b.type(torch.float64)
    assert d.surf_vars["2t"].dtype == d.metadata.lat.dtype == torch.float64
    assert stats.level_to_str(850) == "850" and stats.level_to_str(12.5) == "12_5"


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")
def test_against_reference_classes_live():
    sys.path[:0] = [str(REF), str(Path(__file__).parent / "_shims")]
    try:
        import aurora
        from aurora import normalisation as rn
    finally:
        del sys.path[:2]
    # every normalisation constant of the reference is in our table, bit for bit as float64
    assert stats.locations == {k: float(v) for k, v in rn.locations.items()}
    assert stats.scales == {k: float(v) for k, v in rn.scales.items()}
    cfg = fx.CONFIGS["tiny_air"]
    b = fx.make_batch(cfg, 46, 90, levels=fx.LEVELS13, b=2)
    rb = aurora.Batch(dict(b.surf_vars), dict(b.static_vars), dict(b.atmos_vars),
                      aurora.Metadata(lat=b.metadata.lat, lon=b.metadata.lon, time=b.metadata.time,
                                      atmos_levels=b.metadata.atmos_levels))
    for ours, ref in ((b.normalise(), rb.normalise(surf_stats={})),
                      (b.normalise().unnormalise(), rb.normalise(surf_stats={}).unnormalise(surf_stats={})),
                      (b.crop(3), rb.crop(3))):
        assert ours.spatial_shape == tuple(ref.spatial_shape)
        for grp in ("surf_vars", "static_vars", "atmos_vars"):
            assert list(getattr(ours, grp)) == list(getattr(ref, grp))
            for k, v in getattr(ref, grp).items():
                assert torch.equal(getattr(ours, grp)[k], v), (grp, k)
        assert torch.equal(ours.metadata.lat, ref.metadata.lat)


def test_regrid_properties():
    """Bilinear re-gridding (aurora/batch.py:192-222): a field that is linear in latitude and piecewise linear in
    longitude is reproduced exactly, re-gridding to the same grid is the identity, shapes follow the reference."""
    cfg = fx.CONFIGS["tiny"]
    b = fx.make_batch(cfg, 17, 32, levels=fx.LEVELS4)          # 11.25-degree grid including both poles
    same = b.regrid(11.25)
    assert same.spatial_shape == (17, 32)
    for k in b.surf_vars:
        assert torch.allclose(same.surf_vars[k], b.surf_vars[k], rtol=5e-6, atol=1e-4)
    fine = b.regrid(5.625)
    assert fine.spatial_shape == (33, 64) and fine.atmos_vars["t"].shape == (1, 2, 4, 33, 64)
    assert fine.metadata.lat.dtype == torch.float64 and float(fine.metadata.lat[0]) == 90.0 and float(fine.metadata.lon[-1]) < 360
    lat, lon = b.metadata.lat.double(), b.metadata.lon.double()
    ramp = (3.0 * lat[:, None] + 0.0 * lon[None, :]).float()
    bb = Batch({"2t": ramp[None, None]}, {"z": ramp}, {"t": ramp[None, None, None]}, b.metadata)
    out = bb.regrid(5.625)
    want = (3.0 * out.metadata.lat[:, None] + 0.0 * out.metadata.lon[None, :]).float()
    assert torch.allclose(out.static_vars["z"], want, atol=1e-4)
    # every second point of the finer grid is an original grid point
    assert torch.allclose(fine.surf_vars["2t"][..., ::2, ::2], b.surf_vars["2t"], rtol=5e-6, atol=1e-4)


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("res", [5.0, 11.25, 7.3])
def test_regrid_against_reference_live(res):
    sys.path[:0] = [str(REF), str(Path(__file__).parent / "_shims")]
    try:
        import aurora
    finally:
        del sys.path[:2]
    cfg = fx.CONFIGS["tiny"]
    b = fx.make_batch(cfg, 17, 32, levels=fx.LEVELS4, b=2)
    rb = aurora.Batch(dict(b.surf_vars), dict(b.static_vars), dict(b.atmos_vars),
                      aurora.Metadata(lat=b.metadata.lat, lon=b.metadata.lon, time=b.metadata.time,
                                      atmos_levels=b.metadata.atmos_levels))
    ours, ref = b.regrid(res), rb.regrid(res)
    assert ours.spatial_shape == tuple(ref.spatial_shape)
    assert torch.allclose(ours.metadata.lat, ref.metadata.lat, atol=1e-12) and torch.allclose(ours.metadata.lon, ref.metadata.lon, atol=1e-12)
    for grp in ("surf_vars", "static_vars", "atmos_vars"):
        for k, v in getattr(ref, grp).items():
            got = getattr(ours, grp)[k]
            assert got.dtype == torch.float32 and got.shape == v.shape
            assert torch.allclose(got, v, rtol=2e-6, atol=1e-6 * float(v.abs().max())), (grp, k, float((got - v).abs().max()))


def test_netcdf_io_needs_xarray_like_the_reference(tmp_path):
    try:
        import xarray  # noqa: F401
    except ImportError:
        b = fx.make_batch(fx.CONFIGS["tiny"], 16, 32, levels=fx.LEVELS4)
        with pytest.raises(RuntimeError, match="`xarray` must be installed"):
            b.to_netcdf(tmp_path / "b.nc")
        with pytest.raises(RuntimeError, match="`xarray` must be installed"):
            Batch.from_netcdf(tmp_path / "b.nc")
        return
    b = fx.make_batch(fx.CONFIGS["tiny"], 16, 32, levels=fx.LEVELS4)   # round trip when the libraries are there
    b.to_netcdf(tmp_path / "b.nc")
    r = Batch.from_netcdf(tmp_path / "b.nc")
    for k in b.surf_vars:
        assert torch.equal(r.surf_vars[k], b.surf_vars[k])
    assert r.metadata.time == b.metadata.time and tuple(r.metadata.atmos_levels) == tuple(b.metadata.atmos_levels)
