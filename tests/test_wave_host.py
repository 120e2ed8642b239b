"""Host-side logic of the AuroraWave variant (no GPU): the channel bookkeeping of the engine must replay the
dict mutations of the reference's hooks (aurora.py:851-920) exactly — checked against the oracle, which is
itself pinned to the reference (tests/test_oracle_golden.py), and against the golden files' key order."""

from pathlib import Path

import dataclasses
import numpy as np
import torch

import aurora_b200 as ab
from aurora_b200 import cabi
from aurora_b200.engine import wave_channels, wave_outputs
from oracle import aurora_oracle as O
from tests import fixtures as fx

GOLD = Path(__file__).parent / "golden"


def test_wave_model_declares_reference_channels():
    cfg = fx.CONFIGS["tiny_wave"]
    model = ab.AuroraWave(**fx.our_kwargs(cfg, "AuroraWave"))
    assert dataclasses.replace(model.config, autocast=False) == cfg
    assert model.config.surf_vars == fx.wave_supplemented()
    default = ab.AuroraWave(_init="empty").config  # 1.3 B preset, parameters left uninitialised
    assert default.surf_vars == fx.wave_supplemented() and default.static_vars == fx.WAVE_STATIC
    assert default.lora_mode == "from_second" and default.stabilise_level_agg
    model.load_state_dict(fx.make_state_dict(cfg, seed=8), strict=True)


def test_channel_order_matches_pre_encoder_hook():
    for raw in (fx.WAVE_RAW_SURF, fx.WAVE_RAW_SURF[:-2] + ("dwi",), ("2t", "mwd", "swh", "10u")):
        dummy = {n: torch.zeros(1) for n in raw if n != "dwi"}
        ref_order = list(O._wave_pre(dummy, fx.WAVE_VARS, fx.WAVE_ANGLES))
        ch = wave_channels(tuple(dummy), fx.WAVE_VARS, fx.WAVE_ANGLES)
        assert [k for k, _, _ in ch] == ref_order
        for k, src, tr in ch:
            want = (cabi.AB_IN_DENSITY if k.endswith("_density") else cabi.AB_IN_SIN_DEG if k.endswith("_sin")
                    else cabi.AB_IN_COS_DEG if k.endswith("_cos") else cabi.AB_IN_NAN_TO_ZERO if k in fx.WAVE_VARS
                    else cabi.AB_IN_PLAIN)
            assert tr == want and src == k.removesuffix("_density").removesuffix("_sin").removesuffix("_cos")


def test_output_order_matches_reference_golden():
    gold = np.load(GOLD / "model_tiny_wave_33x64_step1.npz")
    keys = [k[5:] for k in gold.files if k.startswith("surf.")]
    outs = wave_outputs(wave_channels(fx.WAVE_RAW_SURF, fx.WAVE_VARS, fx.WAVE_ANGLES), fx.WAVE_VARS, fx.WAVE_ANGLES)
    assert [k for k, *_ in outs] == keys
    for k, val, cos, dens in outs:
        if k in fx.WAVE_ANGLES:
            assert (val, cos, dens) == (f"{k}_sin", f"{k}_cos", f"{k}_density")
        elif k in fx.WAVE_VARS:
            assert (val, cos, dens) == (k, None, f"{k}_density")
        else:
            assert (val, cos, dens) == (k, None, None)


def test_batch_transform_hook_matches_oracle():
    cfg = fx.CONFIGS["tiny_wave"]
    model = ab.AuroraWave(**fx.our_kwargs(cfg, "AuroraWave"))
    for step, with_dwi in ((0, True), (1, False), (0, False)):
        batch = fx.make_wave_batch(cfg, 33, 64, seed=8, rollout_step=step, with_dwi=with_dwi)
        ours = model.batch_transform_hook(batch)
        ref = O.wave_batch_transform(batch, fx.WAVE_ANGLES)
        assert list(ours.surf_vars) == list(ref.surf_vars) and "dwi" not in ours.surf_vars
        for k in ref.surf_vars:
            assert torch.equal(torch.isnan(ours.surf_vars[k]), torch.isnan(ref.surf_vars[k])), k
            assert torch.equal(ours.surf_vars[k].nan_to_num(0), ref.surf_vars[k].nan_to_num(0)), k
        again = model.batch_transform_hook(ours)  # idempotent (aurora.py:394-399)
        for k in ours.surf_vars:
            assert torch.equal(again.surf_vars[k].nan_to_num(-1), ours.surf_vars[k].nan_to_num(-1)), k
        if step == 0:  # calm patches became NaN, for the whole wave family
            assert torch.isnan(ours.surf_vars["swh"]).sum() > torch.isnan(batch.surf_vars["swh"]).sum()
            assert torch.equal(torch.isnan(ours.surf_vars["swh"]), torch.isnan(ours.surf_vars["mwp"]))
