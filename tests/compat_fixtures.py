"""Synthetic checkpoints in the key layout of the PUBLISHED files (before `_adapt_checkpoint`), tiny dimensions,
deterministic contents.  tests/golden/make_golden.py pushes them through the reference's adapters
(`aurora/model/compat.py`) and stores shape + SHA-256 of every resulting tensor in tests/golden/compat.json;
tests/test_compat.py pushes the same dicts through `aurora_b200.compat` and compares."""

from __future__ import annotations

import hashlib

import numpy as np
import torch

D, T = 8, 2
LEVELS = (50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000)
CAMS_SURF = ("pm1", "pm2p5", "pm10", "tcco", "tc_no", "tcno2", "gtco3", "tcso2")
CAMS_ATMOS = ("co", "no", "no2", "go3", "so2")
PATCH = {"pretrained": 4, "air_pollution": 3, "wave": 4}


def _t(key: str, *shape: int) -> torch.Tensor:
    """Tensor whose every element is distinct and depends on the key."""
    seed = int.from_bytes(hashlib.sha256(key.encode()).digest()[:4], "little")
    n = int(np.prod(shape))
    return torch.from_numpy((np.arange(n, dtype=np.float32) * 1e-3 + seed % 9973).reshape(shape))


def _fill(d: dict, key: str, *shape: int) -> None:
    d[key] = _t(key, *shape)


def old_checkpoint(kind: str) -> dict[str, torch.Tensor]:
    p = PATCH[kind]
    pp = p * p
    d: dict[str, torch.Tensor] = {}
    pre = "net." if kind == "pretrained" else ""
    _fill(d, f"{pre}encoder.surf_token_embeds.weight", D, 7, T, p, p)
    _fill(d, f"{pre}encoder.surf_token_embeds.bias", D)
    _fill(d, f"{pre}encoder.atmos_token_embeds.weight", D, 5, T, p, p)
    _fill(d, f"{pre}encoder.atmos_token_embeds.bias", D)
    _fill(d, f"{pre}decoder.surf_head.weight", 4 * pp, 2 * D)
    _fill(d, f"{pre}decoder.surf_head.bias", 4 * pp)
    for k in ("k_ln", "q_ln") if kind == "wave" else ("ln_k", "ln_q"):
        _fill(d, f"{pre}encoder.level_agg.layers.0.0.{k}.weight", D)
        _fill(d, f"{pre}encoder.level_agg.layers.0.0.{k}.bias", D)
    _fill(d, f"{pre}backbone.encoder_layers.0.blocks.0.attn.qkv.weight", 3 * D, D)
    _fill(d, f"{pre}decoder.level_decoder.layers.0.0.to_q.weight", D, 2 * D)
    if kind != "air_pollution":
        _fill(d, f"{pre}decoder.atmos_head.weight", 5 * pp, 2 * D)
        _fill(d, f"{pre}decoder.atmos_head.bias", 5 * pp)
        return d
    # ---- the CAMS fine-tune's extra generation of modules ----
    _fill(d, "encoder.surf_token_embeds.weight_new", D, 22, T, p, p)
    _fill(d, "encoder.atmos_token_embeds.weight_new", D, 5, T, p, p)
    _fill(d, "encoder.atmos_token_embeds.weight_new2", D, 17, T, p, p)
    for lv in LEVELS:
        _fill(d, f"encoder.atmos_token_embeds_new.layers.{lv}.weight", D, 5, T, p, p)
        _fill(d, f"encoder.atmos_token_embeds_new.layers.{lv}.weight_new", D, 5, T, p, p)
        _fill(d, f"encoder.atmos_token_embeds_new.layers.{lv}.weight_new2", D, 17, T, p, p)
        _fill(d, f"encoder.atmos_token_embeds_new.layers.{lv}.bias", D)
        for sfx, n_old in (("", 5), ("_mod", 5)):
            _fill(d, f"decoder.atmos_head{sfx}.layers.{lv}.weight", n_old * pp, 2 * D)
            _fill(d, f"decoder.atmos_head{sfx}.layers.{lv}.bias", n_old * pp)
            _fill(d, f"decoder.atmos_head{sfx}_new.layers.{lv}.weight", 5 * pp, 2 * D)
            _fill(d, f"decoder.atmos_head{sfx}_new.layers.{lv}.bias", 5 * pp)
    for name in ("2t", "10u", "10v", "msl") + CAMS_SURF:
        _fill(d, f"surf_feature_combiner.{name}.weight", 1, 2)
        _fill(d, f"surf_feature_combiner.{name}.bias", 1)
    for name in ("z", "u", "v", "t", "q") + CAMS_ATMOS:
        _fill(d, f"atmos_feature_combiner.{name}.weight", 1, 2)
        _fill(d, f"atmos_feature_combiner.{name}.bias", 1)
    _fill(d, "decoder.level_decoder_new.layers.0.0.to_q.weight", D, 2 * D)
    _fill(d, "decoder.level_decoder_new.layers.0.1.net.0.bias", D)
    _fill(d, "decoder.surf_head_new.weight", 8 * pp, 2 * D)
    _fill(d, "decoder.surf_head_new.bias", 8 * pp)
    _fill(d, "decoder.surf_head_mod.weight", 12 * pp, 2 * D)
    _fill(d, "decoder.surf_head_mod.bias", 12 * pp)
    return d


def digest(d: dict[str, torch.Tensor]) -> dict[str, list]:
    return {k: [list(v.shape), hashlib.sha256(v.contiguous().numpy().tobytes()).hexdigest()] for k, v in sorted(d.items())}
