"""``Batch`` and ``Metadata``: the data model at the boundary of ``Aurora.forward``.

Same field names, shapes, validation and error behaviour as the reference (`aurora/batch.py:24-190`),
with the methods on the forward path — ``normalise`` / ``unnormalise`` / ``crop`` / ``to`` / ``type`` — and
``regrid`` (`aurora/batch.py:192-222, 299-362`) and netCDF I/O in the reference's file layout
(`aurora/batch.py:224-296`; needs xarray + netCDF4, raises the reference's error without them).  On the GPU path normalisation is fused into the patch-embedding loader and
un-normalisation into the head's store; the methods here exist for API parity and for callers that
want the tensors themselves.
"""

from __future__ import annotations

import dataclasses
from datetime import datetime
from typing import Callable, Optional

import torch

from aurora_b200.stats import atmos_stats_of, surf_stats_of

__all__ = ["Metadata", "Batch"]


@dataclasses.dataclass
class Metadata:
    """Metadata of a batch (`aurora/batch.py:24-68`).

    Args:
        lat: Latitudes, strictly decreasing vector (or matrix), in [-90, 90].
        lon: Longitudes, strictly increasing vector (or matrix), in [0, 360).
        time: One ``datetime`` per batch element.
        atmos_levels: Pressure levels in hPa.
        rollout_step: 0 for analysis data; the model increments it for every prediction.
    """

    lat: torch.Tensor
    lon: torch.Tensor
    time: tuple[datetime, ...]
    atmos_levels: tuple[int | float, ...]
    rollout_step: int = 0

    def __post_init__(self) -> None:
        """Coordinate checks with the reference's messages (`aurora/batch.py:45-68`)."""
        lat, lon = self.lat, self.lon

        def require(ok: torch.Tensor, message: str) -> None:
            if not bool(ok):
                raise ValueError(message)

        require(((lat <= 90) & (lat >= -90)).all(), "Latitudes must be in the range [-90, 90].")
        require(((lon >= 0) & (lon < 360)).all(), "Longitudes must be in the range [0, 360).")
        if lat.dim() == 1 and lon.dim() == 1:
            require((lat.diff() < 0).all(), "Latitudes must be strictly decreasing.")
            require((lon.diff() > 0).all(), "Longitudes must be strictly increasing.")
        elif lat.dim() == 2 and lon.dim() == 2:
            # (the reference only requires non-zero latitude steps for matrix-valued coordinates, batch.py:61)
            require((lat.diff(dim=0) != 0).all(), "Latitudes must be strictly decreasing along every column.")
            require((lon.diff(dim=1) > 0).all(), "Longitudes must be strictly increasing along every row.")
        else:
            raise ValueError("The latitudes and longitudes must either both be vectors or both be matrices.")

    @classmethod
    def derived(cls, lat: torch.Tensor, lon: torch.Tensor, time, atmos_levels, rollout_step: int = 0) -> "Metadata":
        """Metadata whose coordinates come from an already validated instance through a value-preserving step
        (dtype cast, device move, dropping the last latitude row, a latitude band, the next time step): the checks
        of `__post_init__` are not repeated.  On CUDA tensors each of them is a device synchronisation, which
        would stall the launch queue once per forward step."""
        m = object.__new__(cls)
        m.lat, m.lon, m.time, m.atmos_levels, m.rollout_step = lat, lon, time, atmos_levels, rollout_step
        return m


def _affine(x: torch.Tensor, loc, scale, inverse: bool) -> torch.Tensor:
    return x * scale + loc if inverse else (x - loc) / scale


def _norm_surf(x, name, stats, inverse):
    loc, scale = surf_stats_of(name, stats)
    return _affine(x, loc, scale, inverse)


def _norm_atmos(x, name, levels, inverse):
    locs, scs = atmos_stats_of(name, levels)
    loc = torch.tensor(locs, dtype=x.dtype, device=x.device)[..., None, None]
    scale = torch.tensor(scs, dtype=x.dtype, device=x.device)[..., None, None]
    return _affine(x, loc, scale, inverse)


@dataclasses.dataclass
class Batch:
    """A batch of data (`aurora/batch.py:72-190`).

    Args:
        surf_vars: Surface-level variables, each ``(b, t, h, w)``.
        static_vars: Static variables, each ``(h, w)``.
        atmos_vars: Atmospheric variables, each ``(b, t, c, h, w)``.
        metadata: Associated :class:`Metadata`.
    """

    surf_vars: dict[str, torch.Tensor]
    static_vars: dict[str, torch.Tensor]
    atmos_vars: dict[str, torch.Tensor]
    metadata: Metadata

    @property
    def spatial_shape(self) -> tuple[int, int]:
        return tuple(next(iter(self.surf_vars.values())).shape[-2:])

    def _map_vars(self, fs, fst, fa) -> "Batch":
        return Batch(
            surf_vars={k: fs(k, v) for k, v in self.surf_vars.items()},
            static_vars={k: fst(k, v) for k, v in self.static_vars.items()},
            atmos_vars={k: fa(k, v) for k, v in self.atmos_vars.items()},
            metadata=self.metadata,
        )

    def normalise(self, surf_stats: Optional[dict[str, tuple[float, float]]] = None) -> "Batch":
        """``(x - location) / scale`` per variable (and per level for atmospheric variables)."""
        lv = self.metadata.atmos_levels
        return self._map_vars(
            lambda k, v: _norm_surf(v, k, surf_stats, False),
            lambda k, v: _norm_surf(v, k, surf_stats, False),
            lambda k, v: _norm_atmos(v, k, lv, False),
        )

    def unnormalise(self, surf_stats: Optional[dict[str, tuple[float, float]]] = None) -> "Batch":
        """Inverse of :meth:`normalise`."""
        lv = self.metadata.atmos_levels
        return self._map_vars(
            lambda k, v: _norm_surf(v, k, surf_stats, True),
            lambda k, v: _norm_surf(v, k, surf_stats, True),
            lambda k, v: _norm_atmos(v, k, lv, True),
        )

    def crop(self, patch_size: int) -> "Batch":
        """Drop the last latitude row when ``h % patch_size == 1`` (e.g. 721 -> 720)."""
        h, w = self.spatial_shape
        if w % patch_size != 0:
            raise ValueError("Width of the data must be a multiple of the patch size.")
        if h % patch_size == 0:
            return self
        if h % patch_size == 1:
            cut = lambda _k, v: v[..., :-1, :]  # noqa: E731
            out = self._map_vars(cut, cut, cut)
            out.metadata = Metadata.derived(
                lat=self.metadata.lat[:-1],
                lon=self.metadata.lon,
                atmos_levels=self.metadata.atmos_levels,
                time=self.metadata.time,
                rollout_step=self.metadata.rollout_step,
            )
            return out
        raise ValueError(
            f"There can at most be one latitude too many, but there are {h % patch_size} too many."
        )

    def _fmap(self, f: Callable[[torch.Tensor], torch.Tensor]) -> "Batch":
        return Batch(
            surf_vars={k: f(v) for k, v in self.surf_vars.items()},
            static_vars={k: f(v) for k, v in self.static_vars.items()},
            atmos_vars={k: f(v) for k, v in self.atmos_vars.items()},
            metadata=Metadata.derived(
                lat=f(self.metadata.lat),
                lon=f(self.metadata.lon),
                atmos_levels=self.metadata.atmos_levels,
                time=self.metadata.time,
                rollout_step=self.metadata.rollout_step,
            ),
        )

    def to(self, device: str | torch.device) -> "Batch":
        """Move every tensor (incl. lat/lon) to ``device``."""
        return self._fmap(lambda x: x.to(device))

    def type(self, t) -> "Batch":
        """Convert every tensor (incl. lat/lon) to dtype ``t``."""
        return self._fmap(lambda x: x.type(t))

    # -- outside the accelerated path ---------------------------------------------------------
    def regrid(self, res: float) -> "Batch":
        """Bilinear re-gridding to a regular `res`-degree grid that includes both poles (`aurora/batch.py:192-222`):
        `round(180 / res) + 1` latitudes from 90 to -90, `round(360 / res)` longitudes from 0, periodic in longitude,
        linearly extrapolated in latitude.  Computed in float64 and returned as float32, like the reference, but as
        one vectorised gather per variable on the tensors' own device instead of a SciPy loop over fields."""
        n_lat, n_lon = round(180 / res) + 1, round(360 / res)
        dev = self.metadata.lat.device
        lat_new = torch.linspace(90, -90, n_lat, dtype=torch.float64, device=dev)
        lon_new = torch.arange(n_lon, dtype=torch.float64, device=dev) * (360.0 / n_lon)
        plan = _bilinear_plan(self.metadata.lat, self.metadata.lon, lat_new, lon_new)
        return Batch(
            surf_vars={k: _bilinear_apply(v, plan) for k, v in self.surf_vars.items()},
            static_vars={k: _bilinear_apply(v, plan) for k, v in self.static_vars.items()},
            atmos_vars={k: _bilinear_apply(v, plan) for k, v in self.atmos_vars.items()},
            metadata=Metadata(lat=lat_new, lon=lon_new, time=self.metadata.time,
                              atmos_levels=self.metadata.atmos_levels, rollout_step=self.metadata.rollout_step),
        )

    # -- file I/O in the reference's layout (`aurora/batch.py:224-296`); needs xarray + netCDF4 like the reference ----
    _NC_DIMS = {"surf": ("batch", "history", "latitude", "longitude"), "static": ("latitude", "longitude"),
                "atmos": ("batch", "history", "level", "latitude", "longitude")}

    @staticmethod
    def _xarray():
        try:
            import xarray
        except ImportError as e:
            raise RuntimeError("`xarray` must be installed.") from e
        return xarray

    def to_netcdf(self, path) -> None:
        """One data variable per field, named `surf_<name>` / `static_<name>` / `atmos_<name>`, coordinates
        latitude / longitude / time / level / rollout_step."""
        xr = self._xarray()
        groups = {"surf": self.surf_vars, "static": self.static_vars, "atmos": self.atmos_vars}
        data = {f"{grp}_{name}": (self._NC_DIMS[grp], t.detach().cpu().numpy())
                for grp, fields in groups.items() for name, t in fields.items()}
        md = self.metadata
        coords = {"latitude": md.lat.detach().cpu().numpy(), "longitude": md.lon.detach().cpu().numpy(),
                  "time": list(md.time), "level": list(md.atmos_levels), "rollout_step": md.rollout_step}
        xr.Dataset(data, coords=coords).to_netcdf(path)

    @classmethod
    def from_netcdf(cls, path) -> "Batch":
        xr = cls._xarray()
        ds = xr.load_dataset(path, engine="netcdf4")
        fields: dict[str, dict[str, torch.Tensor]] = {"surf": {}, "static": {}, "atmos": {}}
        for key in ds:
            grp, _, name = str(key).partition("_")
            if grp in fields:
                fields[grp][name] = torch.from_numpy(ds[key].values)
        return cls(
            surf_vars=fields["surf"], static_vars=fields["static"], atmos_vars=fields["atmos"],
            metadata=Metadata(
                lat=torch.from_numpy(ds.latitude.values), lon=torch.from_numpy(ds.longitude.values),
                time=tuple(ds.time.values.astype("datetime64[s]").tolist()), atmos_levels=tuple(ds.level.values),
                rollout_step=int(ds.rollout_step.values)),
        )


def _cell_and_weight(grid: torch.Tensor, new: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """For every `new` coordinate: index i of the grid cell [grid[i], grid[i+1]] used for it (clamped to the first /
    last cell, so points outside the grid are extrapolated) and the normalised distance from grid[i].  `grid` is
    strictly monotonic, in either direction."""
    if grid[-1] < grid[0]:
        grid, new = -grid, -new
    i = (torch.searchsorted(grid, new) - 1).clamp(0, grid.numel() - 2)
    return i, (new - grid[i]) / (grid[i + 1] - grid[i])


def _bilinear_plan(lat, lon, lat_new, lon_new):
    lat, lon = lat.double(), lon.double()
    if not bool((lon.diff() > 0).all()):
        raise AssertionError("Longitudes must be strictly increasing.")
    n = lon.numel()
    lon_ext = torch.cat((lon[-1:] - 360, lon, lon[:1] + 360))         # one wrapped column on either side
    col = torch.cat((torch.tensor([n - 1], device=lon.device), torch.arange(n, device=lon.device),
                     torch.tensor([0], device=lon.device)))            # extended column -> source column
    i, wy = _cell_and_weight(lat, lat_new.double())
    j, wx = _cell_and_weight(lon_ext, lon_new.double())
    return i, wy[:, None], col[j], col[j + 1], wx[None, :]


def _bilinear_apply(v: torch.Tensor, plan) -> torch.Tensor:
    i, wy, j0, j1, wx = plan
    v = v.double()
    top, bot = v[..., i, :], v[..., i + 1, :]
    out = (top[..., j0] * (1 - wy) * (1 - wx) + top[..., j1] * (1 - wy) * wx
           + bot[..., j0] * wy * (1 - wx) + bot[..., j1] * wy * wx)
    return out.float()
