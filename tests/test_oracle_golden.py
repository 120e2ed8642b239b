"""Pin the CPU oracle (oracle/aurora_oracle.py) against outputs of the UNMODIFIED reference stored in
tests/golden/model_*.npz (float32 copies of the reference's float64 forward; made by make_golden.py).

Tolerances: the oracle in float64 must agree to float32 storage precision; in float32 it must stay
inside the reference's own fp32-vs-fp64 spread (rel-mean-abs <= 1e-5)."""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import aurora_oracle as O
from tests import fixtures as fx
from tests.golden.cases import MODEL_CASES

GOLD = Path(__file__).parent / "golden"


def _run(name, dtype):
    cfg, sd, batch, variant, vargs, _ = fx.case_inputs(MODEL_CASES[name])
    taps = {}
    with torch.inference_mode():
        pred = O.forward(cfg, sd, batch, dtype=dtype, taps=taps, variant=variant, variant_args=vargs)
    return pred, taps


@pytest.mark.parametrize("name", sorted(MODEL_CASES))
def test_oracle_fp64_matches_reference(name):
    gold = np.load(GOLD / f"model_{name}.npz")
    pred, taps = _run(name, torch.float64)
    assert int(gold["meta.rollout_step"]) == pred.metadata.rollout_step
    assert float(gold["meta.time0"]) == pred.metadata.time[0].timestamp()
    for tap in ("encoder", "backbone"):
        ref = torch.from_numpy(gold[f"tap.{tap}"])
        assert fx.rel_mean_abs(taps[tap], ref) < 2e-7, tap
    for grp, d in (("surf", pred.surf_vars), ("atmos", pred.atmos_vars)):
        keys = [k[len(grp) + 1:] for k in gold.files if k.startswith(grp + ".")]
        assert sorted(keys) == sorted(d.keys())
        if MODEL_CASES[name][1] == "AuroraWave":
            assert list(d.keys()) == keys  # the hooks' dict order (angles re-appended last) is part of the contract
        for k in keys:
            ref = torch.from_numpy(gold[f"{grp}.{k}"])
            assert d[k].shape == ref.shape
            err, nan_mismatch = fx.field_error(d[k], ref, angle=k in fx.WAVE_ANGLES)
            assert err < 2e-7 and nan_mismatch == 0.0, (grp, k, err, nan_mismatch)
            if MODEL_CASES[name][1] == "AuroraWave" and k in fx.WAVE_VARS:
                assert torch.isnan(ref).any() and not torch.isnan(ref).all(), k  # the fixture exercises both sides


@pytest.mark.parametrize("name", ["tiny_33x64", "small_17x32", "tiny_air_46x90", "tiny_wave_33x64"])
def test_oracle_fp32_within_reference_spread(name):
    gold = np.load(GOLD / f"model_{name}.npz")
    pred, _ = _run(name, torch.float32)
    for grp, d in (("surf", pred.surf_vars), ("atmos", pred.atmos_vars)):
        for k, v in d.items():
            err, nan_mismatch = fx.field_error(v, torch.from_numpy(gold[f"{grp}.{k}"]), angle=k in fx.WAVE_ANGLES)
            assert err < 1e-5 and nan_mismatch < 1e-3, (grp, k, err, nan_mismatch)


def test_oracle_rollout_matches_reference():
    gold = np.load(GOLD / "rollout_tiny_lora.npz")
    cfg = fx.CONFIGS["tiny_lora"]
    sd = fx.make_state_dict(cfg, seed=7)
    batch = fx.make_batch(cfg, 33, 64, levels=fx.LEVELS4, b=1, seed=7)
    with torch.inference_mode():
        for i, pred in enumerate(O.rollout(cfg, sd, batch, steps=3, dtype=torch.float64)):
            assert int(gold[f"step{i}.rollout_step"]) == pred.metadata.rollout_step == i + 1
            for k, v in pred.surf_vars.items():
                assert fx.rel_mean_abs(v, torch.from_numpy(gold[f"step{i}.surf.{k}"])) < 1e-6, (i, k)
            for k, v in pred.atmos_vars.items():
                assert fx.rel_mean_abs(v, torch.from_numpy(gold[f"step{i}.atmos.{k}"])) < 1e-6, (i, k)


def test_oracle_rollout_lora_modes_like_reference_test():
    """The reference's tests/test_rollout.py on the oracle: identical weights, one model with a single LoRA for every
    step and one with a separate LoRA per step — the first prediction agrees, later ones do not; time and
    `rollout_step` advance by one model step each."""
    from datetime import timedelta

    cfg1, cfg2 = fx.CONFIGS["tiny_lora"], fx.CONFIGS["tiny_lora_all"]
    sd2 = fx.make_state_dict(cfg2, seed=5)
    sd1 = {k: v for k, v in sd2.items() if ".loras." not in k or ".loras.0." in k}  # same init, only LoRA 0
    assert set(sd1) == {k for k, _, _ in __import__("aurora_b200.spec", fromlist=["param_specs"]).param_specs(cfg1)}
    batch = fx.make_batch(cfg1, 17, 32, levels=fx.LEVELS4, b=1, seed=5)
    steps = 3
    with torch.inference_mode():
        p1 = list(O.rollout(cfg1, sd1, batch, steps=steps))
        p2 = list(O.rollout(cfg2, sd2, batch, steps=steps))
    assert len(p1) == len(p2) == steps
    for i in range(steps):
        want_time = tuple(t + (i + 1) * timedelta(hours=6) for t in batch.metadata.time)
        assert p1[i].metadata.time == p2[i].metadata.time == want_time
        assert p1[i].metadata.rollout_step == p2[i].metadata.rollout_step == i + 1
        same = np.allclose(p1[i].surf_vars["2t"].numpy(), p2[i].surf_vars["2t"].numpy(), rtol=1e-4)
        assert same if i == 0 else not same, i
