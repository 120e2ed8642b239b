"""Checkpoint contract with the reference (no GPU):

* `state_dict` keys and shapes of every preset equal the reference's (tests/golden/keys.json, written by
  make_golden.py from the reference classes), so reference checkpoints load with `strict=True`;
* published-layout checkpoints are rewritten exactly as the reference's `_adapt_checkpoint` does
  (`aurora/model/compat.py`; tests/golden/compat.json holds shape + SHA-256 of every tensor the reference produces);
* history-size adaptation, mirroring the reference's tests/test_checkpoint_adaptation.py."""

import json
from pathlib import Path

import numpy as np
import pytest
import torch

import aurora_b200 as ab
from tests import compat_fixtures as cf

GOLD = Path(__file__).parent / "golden"


@pytest.mark.parametrize("cls_name", ["Aurora", "AuroraPretrained", "AuroraSmallPretrained", "Aurora12hPretrained",
                                      "AuroraHighRes", "AuroraAirPollution", "AuroraWave"])
def test_state_dict_layout_equals_reference(cls_name):
    ref = json.loads((GOLD / "keys.json").read_text())[cls_name]
    model = getattr(ab, cls_name)(_init="empty")
    ours = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert ours.keys() == ref.keys()
    assert ours == ref


@pytest.mark.parametrize("kind,cls_name", [("pretrained", "AuroraSmallPretrained"),
                                           ("air_pollution", "AuroraAirPollution"), ("wave", "AuroraWave")])
def test_published_layout_is_adapted_like_the_reference(kind, cls_name):
    ref = json.loads((GOLD / "compat.json").read_text())[kind]
    cls = getattr(ab, cls_name)
    holder = type("Holder", (), {"patch_size": cf.PATCH[kind]})()  # the adapters only read `self.patch_size`
    adapted = cls._adapt_checkpoint(holder, cf.old_checkpoint(kind))
    assert cf.digest(adapted) == ref
    # a second pass over an already adapted dict changes nothing
    again = cls._adapt_checkpoint(holder, dict(adapted))
    assert cf.digest(again) == ref


def test_current_layout_passes_through():
    model = ab.AuroraSmallPretrained(_init="empty")
    d = {k: torch.zeros(1) for k in model.state_dict()}
    out = model._adapt_checkpoint(dict(d))
    assert out.keys() == d.keys()


# ---- adapt_checkpoint_max_history_size (aurora.py:469-504), cases of the reference's test file ----
def _ckpt():
    g = torch.Generator().manual_seed(0)
    return {"encoder.surf_token_embeds.weights.0": torch.rand((2, 1, 2, 4, 4), generator=g),
            "encoder.atmos_token_embeds.weights.0": torch.rand((2, 1, 2, 4, 4), generator=g)}


class _HistoryOnly:
    adapt_checkpoint_max_history_size = ab.Aurora.adapt_checkpoint_max_history_size

    def __init__(self, n):
        self.max_history_size = n


@pytest.mark.parametrize("size", [4, 5])
def test_adapt_checkpoint_max_history(size):
    ckpt, orig = _ckpt(), _ckpt()
    _HistoryOnly(size).adapt_checkpoint_max_history_size(ckpt)
    for name, w in ckpt.items():
        assert w.shape[2] == size
        np.testing.assert_array_equal(w[:, :, :2].numpy(), orig[name].numpy())
        assert float(w[:, :, 2:].abs().max()) == 0.0


def test_adapt_checkpoint_max_history_fail():
    with pytest.raises(AssertionError):
        _HistoryOnly(1).adapt_checkpoint_max_history_size(_ckpt())


def test_adapt_checkpoint_max_history_twice():
    ckpt, orig = _ckpt(), _ckpt()
    m = _HistoryOnly(4)
    m.adapt_checkpoint_max_history_size(ckpt)
    m.adapt_checkpoint_max_history_size(ckpt)
    for name, w in ckpt.items():
        assert w.shape[2] == 4
        np.testing.assert_array_equal(w[:, :, :2].numpy(), orig[name].numpy())
        assert float(w[:, :, 2:].abs().max()) == 0.0


def test_load_checkpoint_local_roundtrip(tmp_path):
    """A published-layout file goes through torch.load -> adapt -> history extension -> load_state_dict."""
    cfg_kw = dict(embed_dim=64, num_heads=4, encoder_depths=(2, 2), encoder_num_heads=(1, 2), decoder_depths=(2, 2),
                  decoder_num_heads=(2, 1), use_lora=False, max_history_size=3)
    src = ab.Aurora(**{**cfg_kw, "max_history_size": 2}, _init_seed=3)
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    # rebuild the published layout: stacked patch embeddings, fused heads, `net.` prefix
    old = {}
    for k, v in sd.items():
        if "_token_embeds.weights." in k or k.startswith("decoder.surf_heads.") or k.startswith("decoder.atmos_heads."):
            continue
        old["net." + k] = v
    old["net.encoder.surf_token_embeds.weight"] = torch.cat(
        [sd[f"encoder.surf_token_embeds.weights.{n}"] for n in ("2t", "10u", "10v", "msl", "lsm", "z", "slt")], 1)
    old["net.encoder.atmos_token_embeds.weight"] = torch.cat(
        [sd[f"encoder.atmos_token_embeds.weights.{n}"] for n in ("z", "u", "v", "t", "q")], 1)
    for grp, names in (("surf", ("2t", "10u", "10v", "msl")), ("atmos", ("z", "u", "v", "t", "q"))):
        old[f"net.decoder.{grp}_head.weight"] = torch.stack(
            [sd[f"decoder.{grp}_heads.{n}.weight"] for n in names], 1).reshape(-1, 128)
        old[f"net.decoder.{grp}_head.bias"] = torch.stack([sd[f"decoder.{grp}_heads.{n}.bias"] for n in names], 1).reshape(-1)
    path = tmp_path / "old.ckpt"
    torch.save(old, path)
    dst = ab.Aurora(**cfg_kw, _init_seed=4)
    dst.load_checkpoint_local(str(path), strict=True)
    got = dst.state_dict()
    for k, v in sd.items():
        if "_token_embeds.weights." in k:
            assert torch.equal(got[k][:, :, :2], v) and float(got[k][:, :, 2:].abs().max()) == 0.0, k
        else:
            assert torch.equal(got[k], v), k
