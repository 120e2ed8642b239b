"""Location-independent encodings, computed on the host exactly like the reference (float64 sin/cos of
log-spaced wavelengths) and cached by the engine: they depend only on lat/lon, the pressure levels, the
model time step and the batch time stamps, never on the fields.

* Fourier expansion and its five configured ranges     aurora/model/fourier.py:45-126
* spherical polygon area (one constant)                aurora/area.py:12-52
* patch-centre position / patch-area scale encodings   aurora/model/posencoding.py:17-192
"""

from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "fourier_expansion", "pos_scale_encodings", "POS_RANGE", "SCALE_RANGE", "LEAD_RANGE", "LEVELS_RANGE",
    "ABS_TIME_RANGE",
]

_RADIUS_EARTH_KM = 6378137 / 1000


def _min_patch_area(delta: float) -> float:
    """Area (km^2) of the delta x delta degree cell at the pole: the lower wavelength of the scale
    encoding (fourier.py:98-110 via area.py)."""
    pts = [(90.0, 0.0), (90.0, delta), (90.0 - delta, delta), (90.0 - delta, 0.0)]
    pts = pts + [pts[-1]]  # the reference appends the last vertex once more before summing
    n = len(pts)
    acc = 0.0
    for i in range(n):
        lon_lo = math.radians(pts[i][1])
        lat_mid = math.radians(pts[(i + 1) % n][0])
        lon_up = math.radians(pts[(i + 2) % n][1])
        acc += (lon_up - lon_lo) * math.sin(lat_mid)
    return abs(acc * _RADIUS_EARTH_KM * _RADIUS_EARTH_KM / 2)


_DELTA = 0.01
POS_RANGE = (_DELTA, 720.0)
SCALE_RANGE = (_min_patch_area(_DELTA), 4 * np.pi * _RADIUS_EARTH_KM * _RADIUS_EARTH_KM)
LEAD_RANGE = (1 / 60, 24 * 7 * 3)
LEVELS_RANGE = (0.01, 1e5)
ABS_TIME_RANGE = (1.0, 24 * 365.25)


def fourier_expansion(x: torch.Tensor, d: int, rng: tuple[float, float], assert_range: bool = True) -> torch.Tensor:
    """``(..., n) -> (..., n, d)``: sin then cos of ``2 pi x / wavelength`` over ``d/2`` log-spaced
    wavelengths, in float64, returned as float32.  Raises like the reference when a non-zero input
    is outside the configured range."""
    lower, upper = rng
    ax = x.abs()
    in_range = torch.logical_and(lower <= ax, torch.all(ax <= upper))
    if assert_range and not bool(torch.all(torch.logical_or(in_range, x == 0))):
        raise AssertionError(f"The input tensor is not within the configured range `[{lower}, {upper}]`.")
    if d % 2 != 0:
        raise ValueError("The dimensionality must be a multiple of two.")
    x = x.double()
    wl = torch.logspace(math.log10(lower), math.log10(upper), d // 2, base=10, dtype=torch.float64, device=x.device)
    prod = x[..., None] * (2 * np.pi / wl)
    return torch.cat((torch.sin(prod), torch.cos(prod)), dim=-1).float()


def pos_scale_encodings(d: int, lat: torch.Tensor, lon: torch.Tensor, patch: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Position encoding (patch-centre lat | lon, d/2 each) and scale encoding (sqrt of the patch area
    in km, d) of every patch, both ``(L, d)`` float32."""
    if lat.dim() == lon.dim() == 1:
        grid = torch.stack((lat[:, None].expand(-1, lon.numel()), lon[None, :].expand(lat.numel(), -1)), 0)
    elif lat.dim() == lon.dim() == 2:
        grid = torch.stack((lat, lon), 0)
    else:
        raise ValueError(
            f"Latitudes and longitudes must either both be vectors or both be matrices, "
            f"but have dimensionalities {lat.dim()} and {lon.dim()} respectively."
        )
    grid = grid[None].float()
    g_lat, g_lon = grid[:, 0], grid[:, 1]
    k = (patch, patch)
    centre_lat, centre_lon = F.avg_pool2d(g_lat, k), F.avg_pool2d(g_lon, k)
    lat_max, lat_min = F.max_pool2d(g_lat, k), -F.max_pool2d(-g_lat, k)
    lon_max, lon_min = F.max_pool2d(g_lon, k), -F.max_pool2d(-g_lon, k)
    assert (lat_max > lat_min).all() and (lon_max > lon_min).all()
    area = (
        6371**2
        * torch.pi
        * (torch.sin(torch.deg2rad(lat_max)) - torch.sin(torch.deg2rad(lat_min)))
        * (torch.deg2rad(lon_max) - torch.deg2rad(lon_min))
    )
    assert (area > 0.0).all()
    root_area = torch.sqrt(area)
    pos = torch.cat(
        (
            fourier_expansion(centre_lat.reshape(1, -1), d // 2, POS_RANGE),
            fourier_expansion(centre_lon.reshape(1, -1), d // 2, POS_RANGE),
        ),
        dim=-1,
    )[0]
    scale = fourier_expansion(root_area.reshape(1, -1), d, SCALE_RANGE)[0]
    return pos, scale
