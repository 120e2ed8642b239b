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
    cfg_name, cls_name, h, w, levels, bsz, step, seed = MODEL_CASES[name]
    cfg = fx.CONFIGS[cfg_name]
    extra = fx.air_extra_specs(cfg) if cls_name == "AuroraAirPollution" else ()
    sd = fx.make_state_dict(cfg, seed=seed, extra=extra)
    batch = fx.make_batch(cfg, h, w, levels=levels, b=bsz, seed=seed, rollout_step=step)
    taps = {}
    variant = "air_pollution" if cls_name == "AuroraAirPollution" else "base"
    with torch.inference_mode():
        pred = O.forward(cfg, sd, batch, dtype=dtype, taps=taps, variant=variant)
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
        for k in keys:
            ref = torch.from_numpy(gold[f"{grp}.{k}"])
            assert d[k].shape == ref.shape
            err = fx.rel_mean_abs(d[k], ref)
            assert err < 2e-7, (grp, k, err)


@pytest.mark.parametrize("name", ["tiny_33x64", "small_17x32", "tiny_air_46x90"])
def test_oracle_fp32_within_reference_spread(name):
    gold = np.load(GOLD / f"model_{name}.npz")
    pred, _ = _run(name, torch.float32)
    for grp, d in (("surf", pred.surf_vars), ("atmos", pred.atmos_vars)):
        for k, v in d.items():
            err = fx.rel_mean_abs(v, torch.from_numpy(gold[f"{grp}.{k}"]))
            assert err < 1e-5, (grp, k, err)


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
