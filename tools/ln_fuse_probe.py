"""Fused projection + adaLN + residual (ab_gemm_ln_residual) against the unfused pair (ab_gemm_bf16 + ab_ln_mod_residual)
on the Swin shapes of the 0.25-degree model (run on the GPU box)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from aurora_b200 import cabi  # noqa: E402

SHAPES = [("proj_s1", 259200, 512, 512), ("fc2_s1", 259200, 512, 2048), ("proj_s2", 64800, 1024, 1024),
          ("fc2_s2", 64800, 1024, 4096), ("dec_fc2", 842400, 1024, 2048)]


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, m, n, k in SHAPES:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") / k**0.5).to(torch.bfloat16)
    bias, scale, shift = (torch.randn(n, device="cuda") for _ in range(3))
    x = torch.randn(m, n, device="cuda")
    xb = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    y = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)

    def unfused():
        cabi.gemm(a, w, bias=bias, out_bf16=y)
        cabi.ln_mod_residual(y, scale=scale, shift=shift, residual=x, out_f32=x, out_bf16=xb)

    def fused():
        cabi.gemm_ln_residual(a, w, bias=bias, scale=scale, shift=shift, residual=x, out_f32=x, out_bf16=xb)

    t_u, t_f = timed(unfused), timed(fused)
    t_g = timed(lambda: cabi.gemm(a, w, bias=bias, out_bf16=y))
    bytes_f = 2.0 * m * k + 2.0 * n * k + 10.0 * m * n
    print(json.dumps({"name": name, "m": m, "n": n, "k": k, "unfused_ms": round(t_u, 4), "gemm_only_ms": round(t_g, 4),
                      "fused_ms": round(t_f, 4), "fused_tflops": round(2.0 * m * n * k / t_f / 1e9, 1),
                      "fused_gbs": round(bytes_f / t_f / 1e6, 1)}), flush=True)
