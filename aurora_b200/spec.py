"""Model configuration and parameter layout.

``ModelConfig`` carries exactly the constructor arguments of the reference model
(`aurora/model/aurora.py:55-95`) and ``param_specs`` enumerates every ``state_dict`` key with its
shape, so that checkpoints written for the reference load unchanged.  Nothing in
here computes; the CUDA engine (``aurora_b200/engine.py``) and the CPU oracle (``oracle/``) both
consume a plain ``dict[str, Tensor]`` keyed like the reference.
"""

from __future__ import annotations

import dataclasses
import math
from datetime import timedelta
from typing import Iterator, Optional

import torch

from aurora_b200.stats import level_to_str

__all__ = ["ModelConfig", "param_specs", "init_state_dict", "DYNAMIC_VARS"]

DYNAMIC_VARS = ("tod_cos", "tod_sin", "dow_cos", "dow_sin", "doy_cos", "doy_sin")


@dataclasses.dataclass(frozen=True)
class ModelConfig:
    """Hyper-parameters; field names and defaults follow `aurora/model/aurora.py:55-95`."""

    surf_vars: tuple[str, ...] = ("2t", "10u", "10v", "msl")
    static_vars: tuple[str, ...] = ("lsm", "z", "slt")
    atmos_vars: tuple[str, ...] = ("z", "u", "v", "t", "q")
    window_size: tuple[int, int, int] = (2, 6, 12)
    encoder_depths: tuple[int, ...] = (6, 10, 8)
    encoder_num_heads: tuple[int, ...] = (8, 16, 32)
    decoder_depths: tuple[int, ...] = (8, 10, 6)
    decoder_num_heads: tuple[int, ...] = (32, 16, 8)
    latent_levels: int = 4
    patch_size: int = 4
    embed_dim: int = 512
    num_heads: int = 16
    mlp_ratio: float = 4.0
    drop_path: float = 0.0
    drop_rate: float = 0.0
    enc_depth: int = 1
    dec_depth: int = 1
    dec_mlp_ratio: float = 2.0
    perceiver_ln_eps: float = 1e-5
    max_history_size: int = 2
    timestep: timedelta = timedelta(hours=6)
    stabilise_level_agg: bool = False
    use_lora: bool = True
    lora_steps: int = 40
    lora_mode: str = "single"
    surf_stats: Optional[tuple[tuple[str, tuple[float, float]], ...]] = None
    autocast: bool = False
    bf16_mode: bool = False
    level_condition: Optional[tuple[float, ...]] = None
    dynamic_vars: bool = False
    atmos_static_vars: bool = False
    separate_perceiver: tuple[str, ...] = ()
    modulation_heads: tuple[str, ...] = ()
    positive_surf_vars: tuple[str, ...] = ()
    positive_atmos_vars: tuple[str, ...] = ()
    clamp_at_first_step: bool = False
    simulate_indexing_bug: bool = False

    # ---- derived names (mirroring encoder.py:101-114 and decoder.py:78-83) -------------------
    @property
    def enc_static_vars(self) -> tuple[str, ...]:
        sv = tuple(self.static_vars or ())
        if self.dynamic_vars:
            sv = sv + DYNAMIC_VARS
        return sv

    @property
    def enc_surf_embed_vars(self) -> tuple[str, ...]:
        return tuple(self.surf_vars) + self.enc_static_vars

    @property
    def enc_atmos_embed_vars(self) -> tuple[str, ...]:
        av = tuple(self.atmos_vars)
        if self.enc_static_vars and self.atmos_static_vars:
            av = av + tuple(f"static_{v}" for v in self.enc_static_vars)
        return av

    @property
    def dec_surf_vars(self) -> tuple[str, ...]:
        sv = tuple(self.surf_vars)
        return sv + tuple(f"{n}_mod" for n in sv if n in self.modulation_heads)

    @property
    def dec_atmos_vars(self) -> tuple[str, ...]:
        av = tuple(self.atmos_vars)
        return av + tuple(f"{n}_mod" for n in av if n in self.modulation_heads)

    @property
    def dec_separate_perceiver(self) -> tuple[str, ...]:
        sp = tuple(self.separate_perceiver)
        if self.modulation_heads:
            sp = sp + tuple(f"{n}_mod" for n in sp)
        return sp

    @property
    def num_stages(self) -> int:
        return len(self.encoder_depths)


def _perceiver_specs(prefix: str, dim: int, inner: int, depth: int, mlp_ratio: float, ln_k_q: bool):
    hidden = int(dim * mlp_ratio)
    for i in range(depth):
        p = f"{prefix}.layers.{i}"
        yield f"{p}.0.to_q.weight", (inner, dim), "linear_w"
        yield f"{p}.0.to_kv.weight", (2 * inner, dim), "linear_w"
        yield f"{p}.0.to_out.weight", (dim, inner), "linear_w"
        if ln_k_q and i == 0:
            for nm in ("ln_k", "ln_q"):
                yield f"{p}.0.{nm}.weight", (inner,), "ones"
                yield f"{p}.0.{nm}.bias", (inner,), "zeros"
        yield f"{p}.1.net.0.weight", (hidden, dim), "linear_w"
        yield f"{p}.1.net.0.bias", (hidden,), "zeros"
        yield f"{p}.1.net.2.weight", (dim, hidden), "linear_w"
        yield f"{p}.1.net.2.bias", (dim,), "zeros"
        for j in (2, 3):
            yield f"{p}.{j}.weight", (dim,), "ones"
            yield f"{p}.{j}.bias", (dim,), "zeros"


def _patch_embed_specs(prefix: str, var_names, dim: int, t: int, p: int):
    for v in var_names:
        yield f"{prefix}.weights.{v}", (dim, 1, t, p, p), "patch_w"
    yield f"{prefix}.bias", (dim,), "patch_b"


def _block_specs(prefix: str, dim: int, d0: int, mlp_ratio: float, cfg: ModelConfig):
    for nm in ("norm1", "norm2"):
        yield f"{prefix}.{nm}.ln_modulation.1.weight", (2 * dim, d0), "zeros"
        yield f"{prefix}.{nm}.ln_modulation.1.bias", (2 * dim,), "zeros"
    yield f"{prefix}.attn.qkv.weight", (3 * dim, dim), "linear_w"
    yield f"{prefix}.attn.qkv.bias", (3 * dim,), "zeros"
    yield f"{prefix}.attn.proj.weight", (dim, dim), "linear_w"
    yield f"{prefix}.attn.proj.bias", (dim,), "zeros"
    if cfg.use_lora:
        n_lora = cfg.lora_steps if cfg.lora_mode == "all" else 1
        for which, out in (("lora_proj", dim), ("lora_qkv", 3 * dim)):
            for s in range(n_lora):
                yield f"{prefix}.attn.{which}.loras.{s}.lora_A", (8, dim), "lora_a"
                yield f"{prefix}.attn.{which}.loras.{s}.lora_B", (out, 8), "zeros"
    hidden = int(dim * mlp_ratio)
    yield f"{prefix}.mlp.fc1.weight", (hidden, dim), "linear_w"
    yield f"{prefix}.mlp.fc1.bias", (hidden,), "zeros"
    yield f"{prefix}.mlp.fc2.weight", (dim, hidden), "linear_w"
    yield f"{prefix}.mlp.fc2.bias", (dim,), "zeros"


def param_specs(cfg: ModelConfig) -> Iterator[tuple[str, tuple[int, ...], str]]:
    """Yield ``(state_dict key, shape, init kind)`` for every parameter of the model."""
    d0, p, t = cfg.embed_dim, cfg.patch_size, cfg.max_history_size
    # ---- encoder (encoder.py:116-159) ----
    yield "encoder.atmos_latents", (cfg.latent_levels - 1, d0), "latent"
    yield "encoder.surf_level_encoding", (d0,), "latent"
    hid = int(d0 * cfg.mlp_ratio)
    yield "encoder.surf_mlp.net.0.weight", (hid, d0), "linear_w"
    yield "encoder.surf_mlp.net.0.bias", (hid,), "zeros"
    yield "encoder.surf_mlp.net.2.weight", (d0, hid), "linear_w"
    yield "encoder.surf_mlp.net.2.bias", (d0,), "zeros"
    yield "encoder.surf_norm.weight", (d0,), "ones"
    yield "encoder.surf_norm.bias", (d0,), "zeros"
    for nm in ("pos_embed", "scale_embed", "lead_time_embed", "absolute_time_embed", "atmos_levels_embed"):
        yield f"encoder.{nm}.weight", (d0, d0), "linear_w"
        yield f"encoder.{nm}.bias", (d0,), "zeros"
    yield from _patch_embed_specs("encoder.surf_token_embeds", cfg.enc_surf_embed_vars, d0, t, p)
    if not cfg.level_condition:
        yield from _patch_embed_specs("encoder.atmos_token_embeds", cfg.enc_atmos_embed_vars, d0, t, p)
    else:
        for lev in cfg.level_condition:
            yield from _patch_embed_specs(
                f"encoder.atmos_token_embeds.layers.{level_to_str(lev)}", cfg.enc_atmos_embed_vars, d0, t, p
            )
    yield from _perceiver_specs(
        "encoder.level_agg", d0, d0, cfg.enc_depth, cfg.mlp_ratio, cfg.stabilise_level_agg
    )
    # ---- backbone (swin3d.py:805-857) ----
    for i in (0, 2):
        yield f"backbone.time_mlp.{i}.weight", (d0, d0), "linear_w"
        yield f"backbone.time_mlp.{i}.bias", (d0,), "zeros"
    n_enc, n_dec = len(cfg.encoder_depths), len(cfg.decoder_depths)
    for i in range(n_enc):
        dim = d0 * 2**i
        for j in range(cfg.encoder_depths[i]):
            yield from _block_specs(f"backbone.encoder_layers.{i}.blocks.{j}", dim, d0, cfg.mlp_ratio, cfg)
        if i < n_enc - 1:
            yield f"backbone.encoder_layers.{i}.downsample.reduction.weight", (2 * dim, 4 * dim), "linear_w"
            yield f"backbone.encoder_layers.{i}.downsample.norm.weight", (4 * dim,), "ones"
            yield f"backbone.encoder_layers.{i}.downsample.norm.bias", (4 * dim,), "zeros"
    for i in range(n_dec):
        dim = d0 * 2 ** (n_dec - i - 1)
        for j in range(cfg.decoder_depths[i]):
            yield from _block_specs(f"backbone.decoder_layers.{i}.blocks.{j}", dim, d0, cfg.mlp_ratio, cfg)
        if i < n_dec - 1:
            yield f"backbone.decoder_layers.{i}.upsample.lin1.weight", (2 * dim, dim), "linear_w"
            yield f"backbone.decoder_layers.{i}.upsample.lin2.weight", (dim // 2, dim // 2), "linear_w"
            yield f"backbone.decoder_layers.{i}.upsample.norm.weight", (dim // 2,), "ones"
            yield f"backbone.decoder_layers.{i}.upsample.norm.bias", (dim // 2,), "zeros"
    # ---- decoder (decoder.py:93-136) ----
    e = 2 * d0
    yield from _perceiver_specs("decoder.level_decoder", e, e, cfg.dec_depth, cfg.dec_mlp_ratio, False)
    if cfg.dec_separate_perceiver:
        yield from _perceiver_specs(
            "decoder.level_decoder_alternate", e, e, cfg.dec_depth, cfg.dec_mlp_ratio, False
        )
    for v in cfg.dec_surf_vars:
        yield f"decoder.surf_heads.{v}.weight", (p * p, e), "linear_w"
        yield f"decoder.surf_heads.{v}.bias", (p * p,), "zeros"
    for v in cfg.dec_atmos_vars:
        if not cfg.level_condition:
            yield f"decoder.atmos_heads.{v}.weight", (p * p, e), "linear_w"
            yield f"decoder.atmos_heads.{v}.bias", (p * p,), "zeros"
        else:
            for lev in cfg.level_condition:
                yield f"decoder.atmos_heads.{v}.layers.{level_to_str(lev)}.weight", (p * p, e), "linear_w"
                yield f"decoder.atmos_heads.{v}.layers.{level_to_str(lev)}.bias", (p * p,), "zeros"
    yield "decoder.atmos_levels_embed.weight", (e, e), "linear_w"
    yield "decoder.atmos_levels_embed.bias", (e,), "zeros"


def init_state_dict(cfg: ModelConfig, seed: Optional[int] = None, extra=()) -> dict[str, torch.Tensor]:
    """Fresh parameters with the reference's initialisation scheme (`util.py:74-90`, `film.py:34-36`,
    `lora.py:47-51`, `patchembed.py:59-67`, `encoder.py:166-171`): truncated normal (std 0.02) for
    linear weights, zeros for biases / adaLN modulation / LoRA-B, Kaiming-uniform for patch-embedding
    weights and LoRA-A."""
    gen = torch.Generator().manual_seed(seed) if seed is not None else None
    out: dict[str, torch.Tensor] = {}
    for key, shape, kind in list(param_specs(cfg)) + list(extra):
        t = torch.empty(shape, dtype=torch.float32)
        if kind == "zeros":
            t.zero_()
        elif kind == "ones":
            t.fill_(1.0)
        elif kind in ("linear_w", "latent"):
            torch.nn.init.trunc_normal_(t, std=0.02, generator=gen)
        elif kind in ("patch_w", "lora_a"):
            fan_in = math.prod(shape[1:])
            bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform_(a=sqrt(5))
            t.uniform_(-bound, bound, generator=gen)
        elif kind == "patch_b":
            fan_in = cfg.max_history_size * cfg.patch_size**2  # fan-in of ONE per-variable weight
            bound = 1.0 / math.sqrt(fan_in)
            t.uniform_(-bound, bound, generator=gen)
        elif kind == "half":
            t.fill_(0.5)
        else:  # pragma: no cover
            raise ValueError(kind)
        out[key] = t
    return out
