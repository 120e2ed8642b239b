"""The product path has no CPU fallback and no silent degradation (runs without a GPU):
CPU parameters, unsupported parameter dtypes, CPU tensors at the C-ABI wrappers and a missing / unbuildable
library all raise."""

import pytest
import torch

import aurora_b200 as ab
from aurora_b200 import _build, cabi
from tests import fixtures as fx


def _tiny():
    cfg = fx.CONFIGS["tiny"]
    return cfg, ab.Aurora(**fx.our_kwargs(cfg), _init_seed=0)


def test_cpu_parameters_raise():
    cfg, model = _tiny()
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        model.forward(fx.make_batch(cfg, 17, 32, levels=fx.LEVELS4))


def test_double_parameters_raise_not_implemented():
    cfg, model = _tiny()
    with pytest.raises(NotImplementedError, match="fp32 master parameters"):
        model.double().forward(fx.make_batch(cfg, 17, 32, levels=fx.LEVELS4))


def test_fp32_default_is_refused_not_silently_downgraded():
    """`Aurora()` defaults to `autocast=False` = pure fp32 in the reference (aurora.py:84); the engine only has the
    `autocast=True` arithmetic, so `forward` refuses until the caller opts in."""
    cfg = fx.CONFIGS["tiny"]
    model = ab.Aurora(**fx.reference_kwargs(cfg), _init_seed=0)
    assert model.autocast is False
    with pytest.raises(NotImplementedError, match="autocast=True"):
        model.forward(fx.make_batch(cfg, 17, 32, levels=fx.LEVELS4))
    model.autocast = True
    with pytest.raises(RuntimeError, match="CUDA devices only"):  # past the precision gate: now only the device is wrong
        model.forward(fx.make_batch(cfg, 17, 32, levels=fx.LEVELS4))


def test_training_features_raise():
    _, model = _tiny()
    with pytest.raises(NotImplementedError):
        model.configure_activation_checkpointing()


def test_wrappers_reject_cpu_tensors():
    a = torch.zeros(256, 64, dtype=torch.bfloat16)
    w = torch.zeros(64, 64, dtype=torch.bfloat16)
    out = torch.zeros(256, 64, dtype=torch.bfloat16)
    with pytest.raises(cabi.AbError, match="no CPU path"):
        cabi.gemm(a, w, out_bf16=out)


def test_missing_library_is_an_error_not_a_fallback(monkeypatch, tmp_path):
    """No .so and no compiler: loading must raise (nothing else can run the kernels)."""
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(_build, "LIB_PATH", tmp_path / "libaurora_b200.so")
    monkeypatch.setattr(_build, "_nvcc", lambda: (_ for _ in ()).throw(RuntimeError("nvcc not found")))
    with pytest.raises(RuntimeError, match="nvcc not found"):
        cabi.lib()
