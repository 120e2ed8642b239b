"""ORACLE — CPU restatement of the reference's `Aurora.forward` hot path in plain PyTorch tensor ops.

This is TEST INFRASTRUCTURE: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline /
reference arm may import it.  The product (`aurora_b200/`) never does and has no CPU path.

It is a from-scratch functional restatement (no nn.Module, explicit gather maps instead of
roll/pad/partition copies, patch embedding as a GEMM) of:

* Aurora.forward                      aurora/model/aurora.py:265-392 (+ AirPollution hooks :726-796,
                                      AuroraWave hooks :851-920)
* Perceiver3DEncoder.forward          aurora/model/encoder.py:198-366
* LevelPatchEmbed.forward             aurora/model/patchembed.py:79-118
* PerceiverResampler / Attention      aurora/model/perceiver.py:127-152, 212-233
* Swin3DTransformerBackbone.forward   aurora/model/swin3d.py:884-936
* Swin3DTransformerBlock.forward      aurora/model/swin3d.py:440-509, WindowAttention :136-171
* PatchMerging3D / PatchSplitting3D   aurora/model/swin3d.py:526-555, 574-613
* AdaptiveLayerNorm                   aurora/model/film.py:38-49
* LoRA / LoRARollout                  aurora/model/lora.py:53-63, 104-129
* FourierExpansion + instances        aurora/model/fourier.py:45-126, aurora/area.py:12-52
* pos_scale_enc                       aurora/model/posencoding.py:17-192
* Perceiver3DDecoder.forward          aurora/model/decoder.py:168-276, unpatchify aurora/model/util.py:18-41
* rollout                             aurora/rollout.py:14-49

Pinning: checked against the imported reference (randomised zero-init parameters, physically scaled
inputs) by tests/golden/make_golden.py, whose stored outputs tests/test_oracle_golden.py replays.
"""

from __future__ import annotations

import dataclasses
import functools
import math
from datetime import timedelta
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from aurora_b200.batch import Batch, Metadata
from aurora_b200.spec import DYNAMIC_VARS, ModelConfig
from aurora_b200.stats import level_to_str
from oracle import windows as W

Tensor = torch.Tensor

# -------------------------------------------------------------------------------------------------
# Fourier expansions (fourier.py:21-126)
# -------------------------------------------------------------------------------------------------
_RADIUS_EARTH_KM = 6378137 / 1000  # area.py:8


def _polygon_area_km2(poly: Tensor) -> Tensor:
    """Spherical polygon area, lat/lon degrees (area.py:12-52)."""
    poly = torch.cat((poly, poly[..., -1:, :]), dim=-2)
    n = poly.shape[-2]
    acc = torch.zeros(poly.shape[:-2], dtype=poly.dtype)
    if n > 2:
        for i in range(n):
            lo, mid, up = i, (i + 1) % n, (i + 2) % n
            acc = acc + (torch.deg2rad(poly[..., up, 1]) - torch.deg2rad(poly[..., lo, 1])) * torch.sin(
                torch.deg2rad(poly[..., mid, 0])
            )
    return torch.abs(acc * _RADIUS_EARTH_KM * _RADIUS_EARTH_KM / 2)


_DELTA = 0.01
_MIN_PATCH_AREA = _polygon_area_km2(
    torch.tensor([[90, 0], [90, _DELTA], [90 - _DELTA, _DELTA], [90 - _DELTA, 0]], dtype=torch.float64)
).item()
_AREA_EARTH = 4 * np.pi * _RADIUS_EARTH_KM * _RADIUS_EARTH_KM

POS_RANGE = (_DELTA, 720.0)
SCALE_RANGE = (_MIN_PATCH_AREA, _AREA_EARTH)
LEAD_RANGE = (1 / 60, 24 * 7 * 3)
LEVELS_RANGE = (0.01, 1e5)
ABS_TIME_RANGE = (1.0, 24 * 365.25)


def fourier_expansion(x: Tensor, d: int, rng: tuple[float, float], assert_range: bool = True) -> Tensor:
    """sin/cos over d/2 log-spaced wavelengths, computed in float64, returned as float32
    (fourier.py:45-92)."""
    lower, upper = rng
    ax = x.abs()
    in_range = torch.logical_and(lower <= ax, torch.all(ax <= upper))
    if assert_range and not torch.all(torch.logical_or(in_range, x == 0)):
        raise AssertionError(f"The input tensor is not within the configured range `[{lower}, {upper}]`.")
    if d % 2 != 0:
        raise ValueError("The dimensionality must be a multiple of two.")
    x = x.double()
    wl = torch.logspace(math.log10(lower), math.log10(upper), d // 2, base=10, dtype=torch.float64)
    prod = x[..., None] * (2 * np.pi / wl)
    return torch.cat((torch.sin(prod), torch.cos(prod)), dim=-1).float()


def pos_scale_encodings(d: int, lat: Tensor, lon: Tensor, p: int) -> tuple[Tensor, Tensor]:
    """Patch-centre position encoding and patch-root-area scale encoding, both (L, d)
    (posencoding.py:61-192)."""
    if lat.dim() == lon.dim() == 1:
        grid = torch.stack((lat[:, None].expand(-1, lon.numel()), lon[None, :].expand(lat.numel(), -1)), 0)
    elif lat.dim() == lon.dim() == 2:
        grid = torch.stack((lat, lon), 0)
    else:
        raise ValueError("Latitudes and longitudes must either both be vectors or both be matrices.")
    grid = grid[None].float()  # (1, 2, H, W)
    g_lat, g_lon = grid[:, 0], grid[:, 1]
    c_lat = F.avg_pool2d(g_lat, (p, p))
    c_lon = F.avg_pool2d(g_lon, (p, p))
    lat_max, lat_min = F.max_pool2d(g_lat, (p, p)), -F.max_pool2d(-g_lat, (p, p))
    lon_max, lon_min = F.max_pool2d(g_lon, (p, p)), -F.max_pool2d(-g_lon, (p, p))
    area = (
        6371**2
        * torch.pi
        * (torch.sin(torch.deg2rad(lat_max)) - torch.sin(torch.deg2rad(lat_min)))
        * (torch.deg2rad(lon_max) - torch.deg2rad(lon_min))
    )
    assert (area > 0).all()
    root_area = torch.sqrt(area)
    enc_h = fourier_expansion(c_lat.reshape(1, -1), d // 2, POS_RANGE)
    enc_w = fourier_expansion(c_lon.reshape(1, -1), d // 2, POS_RANGE)
    pos = torch.cat((enc_h, enc_w), dim=-1)[0]
    scale = fourier_expansion(root_area.reshape(1, -1), d, SCALE_RANGE)[0]
    return pos, scale


# -------------------------------------------------------------------------------------------------
# small building blocks
# -------------------------------------------------------------------------------------------------
def _lin(sd, key, x, bias=True):
    b = sd.get(f"{key}.bias") if bias else None
    return F.linear(x, sd[f"{key}.weight"].to(x.dtype), None if b is None else b.to(x.dtype))


def _ln(x, w=None, b=None, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), None if w is None else w.to(x.dtype), None if b is None else b.to(x.dtype), eps)


def _mlp(sd, prefix, x):
    """Linear - exact GELU - Linear (perceiver.py:79-84; swin3d.py:59-66)."""
    return _lin(sd, f"{prefix}.2", F.gelu(_lin(sd, f"{prefix}.0", x)))


def patch_embed(sd, prefix: str, x: Tensor, names, p: int) -> Tensor:
    """`LevelPatchEmbed.forward` as a GEMM: x (B, V, T, H, W) -> (B, L, D) with
    K index = ((v*T + t)*P + p1)*P + p2 (patchembed.py:79-118; conv3d with kernel == stride)."""
    b, v, t, h, w = x.shape
    wt = torch.cat([sd[f"{prefix}.weights.{n}"][:, :, :t] for n in names], dim=1).to(x.dtype)  # (D, V, T, P, P)
    d = wt.shape[0]
    cols = x.reshape(b, v, t, h // p, p, w // p, p).permute(0, 3, 5, 1, 2, 4, 6).reshape(b, (h // p) * (w // p), -1)
    return cols @ wt.reshape(d, -1).t() + sd[f"{prefix}.bias"].to(x.dtype)


def perceiver_resampler(sd, prefix, latents, x, num_heads, depth, eps, ln_k_q=False):
    """Post-res-norm Perceiver block(s): latents (R, L1, D), context x (R, L2, D)
    (perceiver.py:127-152, 212-233)."""
    for i in range(depth):
        p = f"{prefix}.layers.{i}"
        q = _lin(sd, f"{p}.0.to_q", latents, bias=False)
        k, v = _lin(sd, f"{p}.0.to_kv", x, bias=False).chunk(2, dim=-1)
        if ln_k_q and i == 0:
            k = _ln(k, sd[f"{p}.0.ln_k.weight"], sd[f"{p}.0.ln_k.bias"])
            q = _ln(q, sd[f"{p}.0.ln_q.weight"], sd[f"{p}.0.ln_q.bias"])
        r, l1, dd = q.shape
        hd = dd // num_heads
        qh = q.reshape(r, l1, num_heads, hd).transpose(1, 2)
        kh = k.reshape(r, -1, num_heads, hd).transpose(1, 2)
        vh = v.reshape(r, -1, num_heads, hd).transpose(1, 2)
        att = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(hd), dim=-1) @ vh
        att = att.transpose(1, 2).reshape(r, l1, dd)
        out = _lin(sd, f"{p}.0.to_out", att, bias=False)
        latents = _ln(out, sd[f"{p}.2.weight"], sd[f"{p}.2.bias"], eps) + latents
        latents = _ln(_mlp(sd, f"{p}.1.net", latents), sd[f"{p}.3.weight"], sd[f"{p}.3.bias"], eps) + latents
    return latents


def _lora_delta(sd, prefix, x, cfg: ModelConfig, step: int):
    """LoRARollout (lora.py:104-129): rank-8 update, alpha/r = 1."""
    if not cfg.use_lora or step >= cfg.lora_steps:
        return 0
    if cfg.lora_mode == "single":
        i = 0
    elif cfg.lora_mode == "from_second":
        if step == 0:
            return 0
        i = 0
    elif cfg.lora_mode == "all":
        i = step
    else:
        raise ValueError(f"Invalid mode: {cfg.lora_mode}")
    a = sd[f"{prefix}.loras.{i}.lora_A"].to(x.dtype)
    b = sd[f"{prefix}.loras.{i}.lora_B"].to(x.dtype)
    return (x @ a.t() @ b.t()) * (8 / 8)


def _ada_ln(sd, prefix, x, c):
    """LN without affine, times scale(c), plus shift(c); shift is the FIRST half (film.py:48-49)."""
    mod = _lin(sd, f"{prefix}.ln_modulation.1", F.silu(c))
    shift, scale = mod.unsqueeze(1).chunk(2, dim=-1)
    return _ln(x) * scale + shift


@functools.lru_cache(maxsize=16)
def _mask_tensor(res, ws0, ss0, dtype):
    """Shifted-window mask as a tensor, memoised per geometry (the reference caches it too, swin3d.py:303)."""
    m = W.shifted_window_mask(res, ws0, ss0, warped=True)
    return None if m is None else torch.from_numpy(np.array(m)).to(dtype)


def swin_block(sd, prefix, x, c, res, num_heads, shifted, cfg: ModelConfig, step: int, taps=None):
    """One Swin3D block (swin3d.py:440-509) via the closed-form gather map of oracle/windows.py."""
    b, l, d = x.shape
    ws0 = tuple(cfg.window_size)
    ss0 = tuple(s // 2 for s in ws0) if shifted else (0, 0, 0)
    idx_np, ws, ss, _ = W.window_gather_map(res, ws0, ss0)
    idx = torch.from_numpy(idx_np)
    nw, n = idx.shape
    flat = idx.reshape(-1)
    valid = flat >= 0
    all_valid = bool(valid.all())
    xw = x.index_select(1, flat.clamp_min(0))
    if not all_valid:
        xw = xw * valid[None, :, None].to(x.dtype)  # zero rows where padded (they still get q = k = v = bias)
    xw = xw.reshape(b, nw, n, d)
    qkv = _lin(sd, f"{prefix}.attn.qkv", xw) + _lora_delta(sd, f"{prefix}.attn.lora_qkv", xw, cfg, step)
    hd = d // num_heads
    qkv = qkv.reshape(b, nw, n, 3, num_heads, hd).permute(3, 0, 1, 4, 2, 5)  # (3, B, nW, H, N, hd)
    q, k, v = qkv[0], qkv[1], qkv[2]
    mask = _mask_tensor(tuple(res), ws0, ss0, x.dtype)
    # softmax(q k^T / sqrt(hd) + mask) v, as the reference's F.scaled_dot_product_attention call (swin3d.py:159-166)
    att = F.scaled_dot_product_attention(
        q, k, v, attn_mask=None if mask is None else mask[None, :, None])
    att = att.permute(0, 1, 3, 2, 4).reshape(b, nw, n, d)
    out = _lin(sd, f"{prefix}.attn.proj", att) + _lora_delta(sd, f"{prefix}.attn.lora_proj", att, cfg, step)
    # reverse partition + crop + un-roll == scatter through the same map (every real token appears exactly once)
    out = out.reshape(b, nw * n, d)
    if all_valid:
        y = torch.empty_like(x).index_copy_(1, flat, out)
    else:
        keep = valid.nonzero().squeeze(1)
        y = torch.zeros_like(x).index_copy_(1, flat[keep], out.index_select(1, keep))
    if taps is not None:
        taps[f"{prefix}.attn_out"] = y
    x = x + _ada_ln(sd, f"{prefix}.norm1", y, c)
    h = _lin(sd, f"{prefix}.mlp.fc2", F.gelu(_lin(sd, f"{prefix}.mlp.fc1", x)))
    x = x + _ada_ln(sd, f"{prefix}.norm2", h, c)
    return x


def patch_merge(sd, prefix, x, res):
    """2x2 spatial gather (zero pad bottom/right to even), LN(4D), Linear 4D->2D (swin3d.py:526-555)."""
    c, h, w = res
    b, l, d = x.shape
    x = x.view(b, c, h, w, d)
    x = F.pad(x, (0, 0, 0, w % 2, 0, h % 2))
    h2, w2 = x.shape[2] // 2, x.shape[3] // 2
    x = x.reshape(b, c, h2, 2, w2, 2, d).permute(0, 1, 2, 4, 3, 5, 6).reshape(b, c * h2 * w2, 4 * d)
    x = _ln(x, sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"])
    return F.linear(x, sd[f"{prefix}.reduction.weight"].to(x.dtype))


def patch_split(sd, prefix, x, res, crop):
    """Linear D->2D, pixel shuffle 2x2, crop the merge padding, LN(D/2), Linear (swin3d.py:574-613)."""
    c, h, w = res
    b, l, d = x.shape
    x = F.linear(x, sd[f"{prefix}.lin1.weight"].to(x.dtype))  # (B, L, 2D)
    x = x.view(b, c, h, w, 2, 2, d // 2).permute(0, 1, 2, 4, 3, 5, 6).reshape(b, c, 2 * h, 2 * w, d // 2)
    # crop_3d with two-sided padding: pad (0, ph, pw) with ph, pw in {0, 1} -> lo = 0, hi = pad.
    ph, pw = crop[1], crop[2]
    x = x[:, :, ph // 2 : 2 * h - (ph - ph // 2), pw // 2 : 2 * w - (pw - pw // 2)]
    x = x.reshape(b, -1, d // 2)
    x = _ln(x, sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"])
    return F.linear(x, sd[f"{prefix}.lin2.weight"].to(x.dtype))


def encoder_specs(patch_res, n_stages):
    """Per-stage resolutions and merge paddings (swin3d.py:868-882)."""
    all_res, padded = [tuple(patch_res)], []
    for _ in range(1, n_stages):
        c, h, w = all_res[-1]
        padded.append((0, h % 2, w % 2))
        all_res.append((c, (h + h % 2) // 2, (w + w % 2) // 2))
    padded.append((0, 0, 0))
    return all_res, padded


def backbone_forward(sd, cfg: ModelConfig, x, patch_res, rollout_step: int, taps=None):
    """3-D Swin U-Net (swin3d.py:884-936)."""
    assert x.shape[1] == patch_res[0] * patch_res[1] * patch_res[2], "Input shape does not match patch size."
    assert patch_res[0] % cfg.window_size[0] == 0
    n_enc, n_dec = len(cfg.encoder_depths), len(cfg.decoder_depths)
    all_res, padded = encoder_specs(patch_res, n_enc)
    hours = cfg.timestep / timedelta(hours=1)
    lead = hours * torch.ones(x.shape[0], dtype=torch.float32)
    c = fourier_expansion(lead, cfg.embed_dim, LEAD_RANGE).to(x.dtype)
    c = _lin(sd, "backbone.time_mlp.2", F.silu(_lin(sd, "backbone.time_mlp.0", c)))
    skips = []
    for i in range(n_enc):
        for j in range(cfg.encoder_depths[i]):
            x = swin_block(
                sd, f"backbone.encoder_layers.{i}.blocks.{j}", x, c, all_res[i], cfg.encoder_num_heads[i],
                j % 2 == 1, cfg, rollout_step, taps,
            )
            if taps is not None:
                taps[f"backbone.encoder_layers.{i}.blocks.{j}"] = x
        skips.append(x)
        if i < n_enc - 1:
            x = patch_merge(sd, f"backbone.encoder_layers.{i}.downsample", x, all_res[i])
    for i in range(n_dec):
        index = n_dec - i - 1
        for j in range(cfg.decoder_depths[i]):
            x = swin_block(
                sd, f"backbone.decoder_layers.{i}.blocks.{j}", x, c, all_res[index], cfg.decoder_num_heads[i],
                j % 2 == 1, cfg, rollout_step, taps,
            )
            if taps is not None:
                taps[f"backbone.decoder_layers.{i}.blocks.{j}"] = x
        if i < n_dec - 1:
            x = patch_split(sd, f"backbone.decoder_layers.{i}.upsample", x, all_res[index], padded[index - 1])
        if 0 < i < n_dec - 1:
            x = x + skips[index - 1]
        elif i == n_dec - 1:
            x = torch.cat([x, skips[0]], dim=-1)
    return x


# -------------------------------------------------------------------------------------------------
# encoder / decoder
# -------------------------------------------------------------------------------------------------
def _dynamic_fields(time, t, h, w, dtype):
    rows = []
    for tm in time:
        vals = (
            np.cos(2 * np.pi * tm.hour / 24), np.sin(2 * np.pi * tm.hour / 24),
            np.cos(2 * np.pi * tm.weekday() / 7), np.sin(2 * np.pi * tm.weekday() / 7),
            np.cos(2 * np.pi * tm.day / 365.25), np.sin(2 * np.pi * tm.day / 365.25),
        )
        ones = torch.ones((1, t, 1, h, w), dtype=dtype)
        rows.append(torch.cat([ones * v for v in vals], dim=-3))
    return torch.cat(rows, dim=0)  # (B, T, 6, H, W)


def encoder_forward(sd, cfg: ModelConfig, batch: Batch, taps=None) -> Tensor:
    """Perceiver3DEncoder.forward (encoder.py:198-366); `batch.static_vars` already (B, T, H, W)."""
    surf_names = tuple(batch.surf_vars)
    static_names = tuple(batch.static_vars)
    atmos_names = tuple(batch.atmos_vars)
    levels = batch.metadata.atmos_levels
    x_surf = torch.stack(tuple(batch.surf_vars.values()), dim=2)
    x_static = torch.stack(tuple(batch.static_vars.values()), dim=2)
    x_atmos = torch.stack(tuple(batch.atmos_vars.values()), dim=2)
    b, t, _, c, h, w = x_atmos.shape
    x_static = x_static.expand((b, t, -1, -1, -1))
    if cfg.dynamic_vars:
        x_dyn = _dynamic_fields(batch.metadata.time, t, h, w, x_static.dtype)
        x_surf = torch.cat((x_surf, x_static, x_dyn), dim=2)
        surf_names = surf_names + static_names + DYNAMIC_VARS
        if cfg.atmos_static_vars:
            atmos_names = atmos_names + tuple(f"static_{v}" for v in static_names + DYNAMIC_VARS)
            ex = (-1, -1, -1, len(levels), -1, -1)
            x_atmos = torch.cat(
                (x_atmos, x_static[..., None, :, :].expand(*ex), x_dyn[..., None, :, :].expand(*ex)), dim=2
            )
    else:
        x_surf = torch.cat((x_surf, x_static), dim=2)
        surf_names = surf_names + static_names
        if cfg.atmos_static_vars:
            atmos_names = atmos_names + static_names
            x_atmos = torch.cat(
                (x_atmos, x_static[..., None, :, :].expand(-1, -1, -1, len(levels), -1, -1)), dim=2
            )
    lat, lon = batch.metadata.lat.float(), batch.metadata.lon.float()
    d, p = cfg.embed_dim, cfg.patch_size

    xs = patch_embed(sd, "encoder.surf_token_embeds", x_surf.transpose(1, 2), surf_names, p)  # (B, L, D)
    dtype = xs.dtype
    if cfg.simulate_indexing_bug and "z" in atmos_names:
        iz, isz = atmos_names.index("z"), atmos_names.index("static_z")
        x_atmos = torch.cat((x_atmos[:, :, :isz], x_atmos[:, :, iz : iz + 1], x_atmos[:, :, isz + 1 :]), dim=2)
    xa_in = x_atmos.permute(0, 3, 2, 1, 4, 5)  # (B, C, V, T, H, W)
    if not cfg.level_condition:
        xa = patch_embed(sd, "encoder.atmos_token_embeds", xa_in.reshape(b * c, *xa_in.shape[2:]), atmos_names, p)
        xa = xa.reshape(b, c, -1, d)
    else:
        xa = torch.stack(
            [
                patch_embed(sd, f"encoder.atmos_token_embeds.layers.{level_to_str(lv)}", xa_in[:, i], atmos_names, p)
                for i, lv in enumerate(levels)
            ],
            dim=1,
        )
    xs = xs + sd["encoder.surf_level_encoding"][None, None, :].to(dtype)
    xs = xs + _ln(_mlp(sd, "encoder.surf_mlp.net", xs), sd["encoder.surf_norm.weight"], sd["encoder.surf_norm.bias"])
    lev_enc = fourier_expansion(torch.tensor(levels), d, LEVELS_RANGE).to(dtype)
    xa = xa + _lin(sd, "encoder.atmos_levels_embed", lev_enc)[None, :, None, :]
    # aggregate_levels (encoder.py:173-196): per location, latents attend over the C levels.
    l = xa.shape[2]
    lat_q = sd["encoder.atmos_latents"].to(dtype)[None].expand(b * l, -1, -1)
    ctx = xa.permute(0, 2, 1, 3).reshape(b * l, c, d)
    agg = perceiver_resampler(
        sd, "encoder.level_agg", lat_q, ctx, cfg.num_heads, cfg.enc_depth, cfg.perceiver_ln_eps,
        ln_k_q=cfg.stabilise_level_agg,
    )
    agg = agg.reshape(b, l, -1, d).permute(0, 2, 1, 3)  # (B, C_latent-1, L, D)
    x = torch.cat((xs.unsqueeze(1), agg), dim=1)
    pos, scale = pos_scale_encodings(d, lat, lon, p)
    x = x + _lin(sd, "encoder.pos_embed", pos[None, None].to(dtype)) + _lin(sd, "encoder.scale_embed", scale[None, None].to(dtype))
    x = x.reshape(b, -1, d)
    hours = cfg.timestep.total_seconds() / 3600
    lead_enc = fourier_expansion(hours * torch.ones(b, dtype=dtype), d, LEAD_RANGE).to(dtype)
    x = x + _lin(sd, "encoder.lead_time_embed", lead_enc).unsqueeze(1)
    abs_t = torch.tensor([tm.timestamp() / 3600 for tm in batch.metadata.time], dtype=torch.float32)
    abs_enc = fourier_expansion(abs_t, d, ABS_TIME_RANGE, assert_range=False)
    x = x + _lin(sd, "encoder.absolute_time_embed", abs_enc.to(dtype)).unsqueeze(1)
    return x


def _unpatchify(x: Tensor, v: int, h: int, w: int, p: int) -> Tensor:
    """(B, L, C, V*P*P) with inner order (P1, P2, V) -> (B, V, C, H, W) (util.py:18-41)."""
    b, c = x.shape[0], x.shape[2]
    hp, wp = h // p, w // p
    x = x.reshape(b, hp, wp, c, p, p, v).permute(0, 6, 3, 1, 4, 2, 5)
    return x.reshape(b, v, c, hp * p, wp * p)


def _head(sd, prefix, x, levels, level_condition):
    if not level_condition:
        return _lin(sd, prefix, x)
    # LevelConditioned along dim -2 (levelcond.py:36-69)
    return torch.stack(
        [_lin(sd, f"{prefix}.layers.{level_to_str(lv)}", x[..., i, :]) for i, lv in enumerate(levels)], dim=-2
    )


def decoder_forward(sd, cfg: ModelConfig, x: Tensor, batch: Batch, patch_res, taps=None) -> Batch:
    """Perceiver3DDecoder.forward (decoder.py:168-276)."""
    surf_names = tuple(batch.surf_vars)
    atmos_names = tuple(batch.atmos_vars)
    levels = batch.metadata.atmos_levels
    surf_names += tuple(f"{n}_mod" for n in surf_names if n in cfg.modulation_heads)
    atmos_names += tuple(f"{n}_mod" for n in atmos_names if n in cfg.modulation_heads)
    b, l, d = x.shape
    lat, lon = batch.metadata.lat.float(), batch.metadata.lon.float()
    h, w = lat.shape[0], lon.shape[-1]
    p = cfg.patch_size
    c0, hp, wp = patch_res
    x = x.reshape(b, c0, hp * wp, d).permute(0, 2, 1, 3)  # (B, HW, C, D)
    xs = torch.stack([_lin(sd, f"decoder.surf_heads.{n}", x[..., :1, :]) for n in surf_names], dim=-1)
    xs = xs.reshape(*xs.shape[:3], -1)
    surf_pred = _unpatchify(xs, len(surf_names), h, w, p).squeeze(2)
    lev_enc = fourier_expansion(torch.tensor(levels), d, LEVELS_RANGE).to(x.dtype)
    lev_emb = _lin(sd, "decoder.atmos_levels_embed", lev_enc)  # (C_A, D)
    q = lev_emb[None].expand(b * hp * wp, -1, -1)
    ctx = x[..., 1:, :].reshape(b * hp * wp, c0 - 1, d)
    heads = cfg.num_heads
    xa = perceiver_resampler(sd, "decoder.level_decoder", q, ctx, heads, cfg.dec_depth, cfg.perceiver_ln_eps)
    xa = xa.reshape(b, hp * wp, len(levels), d)
    if cfg.dec_separate_perceiver:
        xa_alt = perceiver_resampler(
            sd, "decoder.level_decoder_alternate", q, ctx, heads, cfg.dec_depth, cfg.perceiver_ln_eps
        ).reshape(b, hp * wp, len(levels), d)
    else:
        xa_alt = xa
    outs = [
        _head(sd, f"decoder.atmos_heads.{n}", xa_alt if n in cfg.dec_separate_perceiver else xa, levels, cfg.level_condition)
        for n in atmos_names
    ]
    xo = torch.stack(outs, dim=-1)
    xo = xo.reshape(*xo.shape[:3], -1)
    atmos_pred = _unpatchify(xo, len(atmos_names), h, w, p)
    return Batch(
        {v: surf_pred[:, i] for i, v in enumerate(surf_names)},
        batch.static_vars,
        {v: atmos_pred[:, i] for i, v in enumerate(atmos_names)},
        Metadata(
            lat=lat,
            lon=lon,
            time=tuple(tm + cfg.timestep for tm in batch.metadata.time),
            atmos_levels=levels,
            rollout_step=batch.metadata.rollout_step + 1,
        ),
    )


# -------------------------------------------------------------------------------------------------
# AirPollution hooks (aurora.py:726-796)
# -------------------------------------------------------------------------------------------------
_DIFF_DIM = {"pm1": 0, "pm2p5": 0, "pm10": 0, "co": 1, "tcco": 1, "no": 0, "tc_no": 0, "no2": 0, "tcno2": 0,
             "so2": 1, "tcso2": 1, "go3": 1, "gtco3": 1}


def _combine(sd, group, name, z):
    eps = 1e-4
    feats = torch.stack([z.clamp(min=0, max=2.5), (torch.log(z.clamp(min=eps)) - np.log(eps)) / (-np.log(eps))], -1)
    wgt = sd[f"{group}.{name}.weight"].to(z.dtype)
    bias = sd[f"{group}.{name}.bias"].to(z.dtype)
    return F.linear(feats, wgt, bias)[..., 0]


def wave_batch_transform(batch: Batch, angle_vars) -> Batch:
    """AuroraWave.batch_transform_hook (aurora.py:851-890): wind speed / direction -> components; at roll-out
    step 0 a wave family whose height is < 1e-4 is marked absent (NaN)."""
    surf = dict(batch.surf_vars)
    if "dwi" in surf and "wind" in surf:
        surf["10u_wave"] = -surf["wind"] * torch.sin(torch.deg2rad(surf["dwi"]))
        surf["10v_wave"] = -surf["wind"] * torch.cos(torch.deg2rad(surf["dwi"]))
        del surf["dwi"]
    if batch.metadata.rollout_step == 0:
        for height, others in (("swh", ("mwd", "mwp", "pp1d")), ("shww", ("mdww", "mpww")),
                               ("shts", ("mdts", "mdts")), ("swh1", ("mwd1", "mwp1")), ("swh2", ("mwd2", "mwp2"))):
            absent = surf[height] < 1e-4
            if absent.sum() > 0:
                for name in (height,) + others:
                    x = surf[name].clone()
                    x[absent] = float("nan")
                    surf[name] = x
    return dataclasses.replace(batch, surf_vars=surf)


def _wave_pre(surf: dict, density_vars, angle_vars) -> dict:
    """AuroraWave._pre_encoder_hook (aurora.py:874-892) on normalised fields: new channels are appended to the
    dict in the order the loop meets them, the angle itself is removed."""
    surf = dict(surf)
    for name in list(surf):
        x = surf[name]
        if name in density_vars and f"{name}_density" not in surf:
            surf[f"{name}_density"] = (~torch.isnan(x)).to(x.dtype)
            surf[name] = x.nan_to_num(0)
        if name in angle_vars and not (f"{name}_sin" in surf and f"{name}_cos" in surf):
            surf[f"{name}_sin"] = torch.sin(torch.deg2rad(x)).nan_to_num(0)
            surf[f"{name}_cos"] = torch.cos(torch.deg2rad(x)).nan_to_num(0)
            del surf[name]
    return surf


def _wave_post(pred_surf: dict, wmb: Tensor, density_vars, angle_vars) -> dict:
    """AuroraWave._post_decoder_hook (aurora.py:894-920), normalised units; `wmb` is the normalised mask."""
    out = dict(pred_surf)
    mask = wmb > 0
    for name in angle_vars:
        if f"{name}_sin" in out and f"{name}_cos" in out:
            out[name] = torch.rad2deg(torch.atan2(out.pop(f"{name}_sin"), out.pop(f"{name}_cos"))) % 360
    for name in density_vars:
        if name in out:
            density = torch.sigmoid(out.pop(f"{name}_density")) * mask
            data = out[name] * mask
            data[density < 0.5] = float("nan")
            out[name] = data
    return out


# -------------------------------------------------------------------------------------------------
# whole model
# -------------------------------------------------------------------------------------------------
def forward(cfg: ModelConfig, sd: dict[str, Tensor], batch: Batch, dtype=torch.float32, taps: Optional[dict] = None,
            variant: str = "base", variant_args: Optional[dict] = None) -> Batch:
    """`Aurora.forward` (aurora.py:265-392) on the CPU.  `variant="air_pollution"` adds the
    AuroraAirPollution pre/post hooks, `variant="wave"` the AuroraWave ones (`variant_args` =
    {"density_vars": ..., "angle_vars": ...})."""
    va = variant_args or {}
    if variant == "wave":
        batch = wave_batch_transform(batch, va["angle_vars"])
    sd = {k: v.to(dtype) for k, v in sd.items()}
    surf_stats = dict(cfg.surf_stats) if cfg.surf_stats else None
    batch = batch.type(dtype)
    batch = batch.normalise(surf_stats=surf_stats)
    batch = batch.crop(patch_size=cfg.patch_size)
    batch = batch.to("cpu")
    h, w = batch.spatial_shape
    patch_res = (cfg.latent_levels, h // cfg.patch_size, w // cfg.patch_size)
    b, t = next(iter(batch.surf_vars.values())).shape[:2]
    batch = dataclasses.replace(batch, static_vars={k: v[None, None].repeat(b, t, 1, 1) for k, v in batch.static_vars.items()})
    tb = batch
    if cfg.positive_surf_vars:
        tb = dataclasses.replace(tb, surf_vars={k: v.clamp(min=0) if k in cfg.positive_surf_vars else v for k, v in batch.surf_vars.items()})
    if cfg.positive_atmos_vars:
        tb = dataclasses.replace(tb, atmos_vars={k: v.clamp(min=0) if k in cfg.positive_atmos_vars else v for k, v in batch.atmos_vars.items()})
    if variant == "air_pollution":
        tb = dataclasses.replace(
            tb,
            surf_vars={k: _combine(sd, "surf_feature_combiner", k, v) if k in cfg.positive_surf_vars else v for k, v in tb.surf_vars.items()},
            atmos_vars={k: _combine(sd, "atmos_feature_combiner", k, v) if k in cfg.positive_atmos_vars else v for k, v in tb.atmos_vars.items()},
        )
    if variant == "wave":
        # The reference's hook MUTATES the surf_vars dict that `batch` and the transformed batch share (no
        # positive variables => same object, aurora.py:299-320), so the decoder sees the derived channels too.
        assert not cfg.positive_surf_vars
        tb = dataclasses.replace(tb, surf_vars=_wave_pre(tb.surf_vars, va["density_vars"], va["angle_vars"]))
        batch = dataclasses.replace(batch, surf_vars=tb.surf_vars)
    x = encoder_forward(sd, cfg, tb, taps)
    if taps is not None:
        taps["encoder"] = x
    x = backbone_forward(sd, cfg, x, patch_res, batch.metadata.rollout_step, taps)
    if taps is not None:
        taps["backbone"] = x
    pred = decoder_forward(sd, cfg, x, batch, patch_res, taps)
    pred = dataclasses.replace(pred, static_vars={k: v[0, 0] for k, v in batch.static_vars.items()})
    pred = dataclasses.replace(
        pred,
        surf_vars={k: v[:, None] for k, v in pred.surf_vars.items()},
        atmos_vars={k: v[:, None] for k, v in pred.atmos_vars.items()},
    )
    if variant == "air_pollution":
        def diff(prev, model, name):
            if name in _DIFF_DIM:
                return model[name] + (1 + model[f"{name}_mod"]) * prev[name][:, _DIFF_DIM[name]]
            return model[name]
        pred = dataclasses.replace(
            pred,
            surf_vars={k: diff(batch.surf_vars, pred.surf_vars, k) for k in batch.surf_vars},
            atmos_vars={k: diff(batch.atmos_vars, pred.atmos_vars, k) for k in batch.atmos_vars},
        )
        if cfg.use_lora:
            parts = []
            for i, lv in enumerate(pred.metadata.atmos_levels):
                sec = pred.atmos_vars["so2"][..., i, :, :]
                parts.append(sec.clamp(max=1) if lv >= 850 else sec)
            pred.atmos_vars["so2"] = torch.stack(parts, dim=-3)
    if variant == "wave":
        pred = dataclasses.replace(pred, surf_vars=_wave_post(
            pred.surf_vars, pred.static_vars["wmb"], va["density_vars"], va["angle_vars"]))
    step = pred.metadata.rollout_step
    clamp_now = step >= 1 if cfg.clamp_at_first_step else step > 1
    if cfg.positive_surf_vars and clamp_now:
        pred = dataclasses.replace(pred, surf_vars={k: v.clamp(min=0) if k in cfg.positive_surf_vars else v for k, v in pred.surf_vars.items()})
    if cfg.positive_atmos_vars and clamp_now:
        pred = dataclasses.replace(pred, atmos_vars={k: v.clamp(min=0) if k in cfg.positive_atmos_vars else v for k, v in pred.atmos_vars.items()})
    return pred.unnormalise(surf_stats=surf_stats)


def rollout(cfg: ModelConfig, sd, batch: Batch, steps: int, dtype=torch.float32, variant="base", variant_args=None):
    """Autoregressive roll-out generator (rollout.py:14-49)."""
    if variant == "wave":
        batch = wave_batch_transform(batch, variant_args["angle_vars"])
    batch = batch.type(dtype).crop(cfg.patch_size).to("cpu")
    for _ in range(steps):
        pred = forward(cfg, sd, batch, dtype=dtype, variant=variant, variant_args=variant_args)
        yield pred
        batch = dataclasses.replace(
            pred,
            surf_vars={k: torch.cat([batch.surf_vars[k][:, 1:], v], dim=1) for k, v in pred.surf_vars.items()},
            atmos_vars={k: torch.cat([batch.atmos_vars[k][:, 1:], v], dim=1) for k, v in pred.atmos_vars.items()},
        )
