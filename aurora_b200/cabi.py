"""ctypes binding of ``libaurora_b200.so`` (the C ABI declared in ``include/aurora_b200.h``).

PyTorch owns every buffer; this module only forwards raw device pointers, sizes and the current
CUDA stream.  There is no CPU fallback: if the library is missing it is built (``nvcc`` required),
and if it cannot be loaded every call raises.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from aurora_b200 import _build

__all__ = ["lib", "AbError", "check", "ptr", "stream_ptr", "gemm", "launch_count"]

AB_ACT_NONE = 0
AB_ACT_GELU_ERF = 1


class AbError(RuntimeError):
    """A call into libaurora_b200.so returned a negative status."""


class AbGemm(C.Structure):
    _fields_ = [
        ("a", C.c_void_p),
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("out_bf16", C.c_void_p),
        ("m", C.c_int32),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("lda", C.c_int32),
        ("ldw", C.c_int32),
        ("ldr", C.c_int32),
        ("ld_f32", C.c_int32),
        ("ld_bf16", C.c_int32),
        ("act", C.c_int32),
    ]


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load (building first if needed) the shared library.  Raises if that is impossible."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        if not path.exists():
            _build.build()
        handle = C.CDLL(str(path))
        handle.ab_version.restype = C.c_int
        handle.ab_last_error.restype = C.c_char_p
        handle.ab_launch_count.restype = C.c_ulonglong
        for name in EXPORTS:
            if name in ("ab_version", "ab_last_error", "ab_launch_count"):
                continue
            getattr(handle, name).restype = C.c_int
        _lib = handle
    return _lib


# Every symbol include/aurora_b200.h declares (tests check that the library exports all of them).
EXPORTS = [
    "ab_version",
    "ab_last_error",
    "ab_launch_count",
    "ab_gemm_bf16",
]


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().ab_last_error().decode("utf-8", "replace")
        raise AbError(f"{what or 'libaurora_b200'} failed with status {status}: {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a CUDA tensor (None passes through as NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise AbError("libaurora_b200 operates on CUDA tensors only (got a CPU tensor); there is no CPU path")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().ab_launch_count())


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D view"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gemm(
    a: torch.Tensor,
    w: torch.Tensor,
    *,
    bias: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    out_f32: Optional[torch.Tensor] = None,
    out_bf16: Optional[torch.Tensor] = None,
    act: int = AB_ACT_NONE,
) -> None:
    """``out = act(a @ w.T + bias) + residual`` on the tcgen05 GEMM (a, w bf16; outputs preallocated)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    m, k = a.shape
    n, k2 = w.shape
    assert k == k2, f"K mismatch {k} vs {k2}"
    g = AbGemm()
    g.a, g.w = ptr(a), ptr(w)
    g.bias = ptr(bias)
    g.residual = ptr(residual)
    g.out_f32 = ptr(out_f32)
    g.out_bf16 = ptr(out_bf16)
    g.m, g.n, g.k = m, n, k
    g.lda, g.ldw = _ld(a), _ld(w)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n and bias.is_contiguous()
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == (m, n)
        g.ldr = _ld(residual)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.shape == (m, n)
        g.ld_f32 = _ld(out_f32)
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16 and out_bf16.shape == (m, n)
        g.ld_bf16 = _ld(out_bf16)
    g.act = act
    check(lib().ab_gemm_bf16(C.byref(g), C.c_void_p(stream_ptr())), "ab_gemm_bf16")
