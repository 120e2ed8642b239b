"""Key layout adaptation of the published checkpoint files.

The published `.ckpt` files predate the name-based parametrisation: variables are indexed by position inside
stacked tensors, heads are fused, the air-pollution fine-tune carries a second set of `*_new` modules and the
wave fine-tune spells two LayerNorm names differently.  The reference rewrites such a dict at load time
(`aurora/model/compat.py:19-78` pretrained, `:81-267` air pollution, `:270-284` wave; called from
`Aurora.load_checkpoint_local`, `aurora.py:432-467`).  This module does the same job as a set of declarative
rewrite rules, so `load_checkpoint_local` accepts the same files as the reference.  Every function is a no-op on
a dict that is already in the current layout (with the one exception the reference has too: the air-pollution
`z` patch embedding is always tied to `static_z`, `compat.py:152-155`).

Host-side loader logic only; nothing here touches the GPU.
"""

from __future__ import annotations

import torch

from aurora_b200.stats import level_to_str

__all__ = ["adapt_pretrained", "adapt_air_pollution", "adapt_wave"]

Tensors = dict[str, torch.Tensor]

ERA5_SURF = ("2t", "10u", "10v", "msl")
ERA5_STATIC = ("lsm", "z", "slt")
ERA5_ATMOS = ("z", "u", "v", "t", "q")
CAMS_SURF = ("pm1", "pm2p5", "pm10", "tcco", "tc_no", "tcno2", "gtco3", "tcso2")
CAMS_ATMOS = ("co", "no", "no2", "go3", "so2")
CAMS_STATIC = ("static_ammonia", "static_ammonia_log", "static_co", "static_co_log", "static_nox", "static_nox_log",
               "static_so2", "static_so2_log")
CLOCK = ("tod_cos", "tod_sin", "dow_cos", "dow_sin", "doy_cos", "doy_sin")
CAMS_LEVELS = (50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000)


def _unstack_channels(d: Tensors, stacked: str, names: tuple[str, ...], dest: str) -> None:
    """`stacked` (D, V, T, P, P), one input channel per variable -> `dest.format(name)` (D, 1, T, P, P) each."""
    if stacked not in d:
        return
    w = d.pop(stacked)
    assert w.shape[1] == len(names)
    for i, name in enumerate(names):
        d[dest.format(name)] = w[:, [i]]


def _unfuse_head(d: Tensors, fused: str, names: tuple[str, ...], patch_size: int, dest: str, keep=None) -> None:
    """Fused head `fused.weight` (P*P*V, D) / `fused.bias` (P*P*V,) with the variable as the FAST index of the
    output dimension -> one (P*P, D) / (P*P,) head per variable under `dest.format(name)`; variables outside
    `keep` (when given) are dropped."""
    if f"{fused}.weight" not in d:
        return
    w, b = d.pop(f"{fused}.weight"), d.pop(f"{fused}.bias")
    if keep is not None and not keep:
        return
    pp, v = patch_size**2, len(names)
    assert w.shape[0] == v * pp
    assert b.shape[0] == v * pp
    w, b = w.reshape(pp, v, -1), b.reshape(pp, v)
    for i, name in enumerate(names):
        if keep is None or name in keep:
            d[dest.format(name) + ".weight"] = w[:, i]
            d[dest.format(name) + ".bias"] = b[:, i]


def _rename(d: Tensors, old: str, new: str, prefix_only: bool) -> None:
    for k in list(d):
        if (k.startswith(old) if prefix_only else old in k):
            d[(new + k[len(old):]) if prefix_only else k.replace(old, new)] = d.pop(k)


def adapt_pretrained(patch_size: int, d: Tensors) -> Tensors:
    """Lightning `net.` prefix, index-based patch embeddings and fused heads (`compat.py:19-78`)."""
    _rename(d, "net.", "", prefix_only=True)
    _unstack_channels(d, "encoder.surf_token_embeds.weight", ERA5_SURF + ERA5_STATIC,
                      "encoder.surf_token_embeds.weights.{}")
    _unstack_channels(d, "encoder.atmos_token_embeds.weight", ERA5_ATMOS, "encoder.atmos_token_embeds.weights.{}")
    _unfuse_head(d, "decoder.surf_head", ERA5_SURF, patch_size, "decoder.surf_heads.{}")
    _unfuse_head(d, "decoder.atmos_head", ERA5_ATMOS, patch_size, "decoder.atmos_heads.{}")
    return d


def adapt_air_pollution(patch_size: int, d: Tensors) -> Tensors:
    """The CAMS fine-tune's second generation of modules (`*_new`, `*_mod`) folded into the level-conditioned,
    name-based layout (`compat.py:81-267`)."""
    emb, emb_new = "encoder.atmos_token_embeds", "encoder.atmos_token_embeds_new"
    _unstack_channels(d, "encoder.surf_token_embeds.weight_new", CAMS_SURF + CAMS_STATIC + CLOCK,
                      "encoder.surf_token_embeds.weights.{}")

    # one shared ERA5 atmospheric embedding -> a copy per pressure level (`compat.py:104-119`)
    if f"{emb}.weights.z" in d and f"{emb_new}.layers.50.weight" in d:
        shared_bias = d.pop(f"{emb}.bias")
        shared = {name: d.pop(f"{emb}.weights.{name}") for name in ERA5_ATMOS}
        for level in CAMS_LEVELS:
            for name, w in shared.items():
                d[f"{emb}.layers.{level}.weights.{name}"] = w.clone()
            d[f"{emb}.layers.{level}.bias"] = shared_bias.clone()

    # static / clock channels of the atmospheric embedding, shared by all levels (`compat.py:121-140`)
    static_names = tuple(f"static_{n}" for n in ERA5_STATIC + CAMS_STATIC + CLOCK)
    if f"{emb}.weight_new2" in d:
        w = d[f"{emb}.weight_new2"]
        assert w.shape[1] == len(static_names)
        for level in CAMS_LEVELS:
            for i, name in enumerate(static_names):
                d[f"{emb}.layers.{level_to_str(level)}.weights.{name}"] = w[:, [i]]
    for stale in (f"{emb}.weight_new", f"{emb}.weight_new2"):
        d.pop(stale, None)

    for level in CAMS_LEVELS:
        lv = level_to_str(level)
        d.pop(f"{emb_new}.layers.{lv}.weight", None)  # doubly specified; the `weight_new` copy is the live one
        _unstack_channels(d, f"{emb_new}.layers.{lv}.weight_new", CAMS_ATMOS, f"{emb}.layers.{lv}.weights.{{}}")
        # `z` uses the patch embedding of `static_z`: the indexing bug the model was trained with (`compat.py:162-165`)
        d[f"{emb}.layers.{lv}.weights.z"] = d[f"{emb}.layers.{lv}.weights.static_z"]
        if f"{emb_new}.layers.{lv}.bias" in d:
            # two patch-embedding instances used to be summed, so their biases add (`compat.py:167-174`)
            assert f"{emb}.layers.{lv}.bias" in d
            d[f"{emb}.layers.{lv}.bias"] += d.pop(f"{emb_new}.layers.{lv}.bias")
        d.pop(f"{emb_new}.layers.{lv}.weight_new2", None)

    # feature combiners exist for the pollution variables only (`compat.py:178-188`)
    for group, names in (("surf_feature_combiner", ERA5_SURF), ("atmos_feature_combiner", ERA5_ATMOS)):
        for name in names:
            if f"{group}.{name}.weight" in d:
                del d[f"{group}.{name}.weight"], d[f"{group}.{name}.bias"]

    _rename(d, "decoder.level_decoder_new", "decoder.level_decoder_alternate", prefix_only=True)

    _unfuse_head(d, "decoder.surf_head_new", CAMS_SURF, patch_size, "decoder.surf_heads.{}")
    _unfuse_head(d, "decoder.surf_head_mod", ERA5_SURF + CAMS_SURF, patch_size, "decoder.surf_heads.{}_mod",
                 keep=CAMS_SURF)
    for suffix in ("", "_mod"):
        for level in CAMS_LEVELS:
            # modulation heads are kept for the pollution variables only (`compat.py:236-250`)
            _unfuse_head(d, f"decoder.atmos_head{suffix}.layers.{level}", ERA5_ATMOS, patch_size,
                         f"decoder.atmos_heads.{{}}{suffix}.layers.{level}", keep=() if suffix == "_mod" else None)
            _unfuse_head(d, f"decoder.atmos_head{suffix}_new.layers.{level}", CAMS_ATMOS, patch_size,
                         f"decoder.atmos_heads.{{}}{suffix}.layers.{level}")
    return d


def adapt_wave(patch_size: int, d: Tensors) -> Tensors:
    """The level-aggregation LayerNorms are spelled `k_ln` / `q_ln` in the wave checkpoint (`compat.py:270-284`)."""
    del patch_size
    for old, new in ((".k_ln.", ".ln_k."), (".q_ln.", ".ln_q.")):
        _rename(d, old, new, prefix_only=False)
    return d
