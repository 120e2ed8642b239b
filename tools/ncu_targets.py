"""One launch each of the kernels that get an `ncu --set full` capture in round 2 (run under ncu on the GPU box):
window attention (P in TMEM) on the stage-1 and stage-3 grids, shifted; the pair GEMM on qkv_s1 / fc1_s1 (GELU) / fc2_s2;
the LN-fused projection on proj_s1; the row kernel on the stage-1 stream; a self-neighbour halo push."""
import sys

import torch

sys.path.insert(0, ".")
from aurora_b200 import cabi  # noqa: E402

WS = (2, 6, 12)
for res, heads in [((4, 180, 360), 8), ((4, 45, 90), 32)]:
    l, d = res[0] * res[1] * res[2], heads * 64
    qkv = torch.randn(l, 3 * d, device="cuda").to(torch.bfloat16)
    pad = torch.randn(3 * d, device="cuda").to(torch.bfloat16)
    o = torch.empty(l, d, device="cuda", dtype=torch.bfloat16)
    cabi.window_attention(qkv, o, batch=1, res=res, window=WS, shift=(1, 3, 6), num_heads=heads, pad_qkv=pad)
    torch.cuda.synchronize()
    del qkv, o
for m, n, k, act in [(259200, 1536, 512, 0), (259200, 2048, 512, 1), (64800, 1024, 4096, 0)]:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") / k**0.5).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    cabi.gemm(a, w, bias=bias, out_bf16=out, act=act)
    torch.cuda.synchronize()
    del a, w, out
m, n, k = 259200, 512, 512
a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
w = (torch.randn(n, k, device="cuda") / k**0.5).to(torch.bfloat16)
bias, scale, shift = (torch.randn(n, device="cuda") for _ in range(3))
x = torch.randn(m, n, device="cuda")
xb = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
y = torch.randn(m, n, device="cuda").to(torch.bfloat16)
cabi.gemm_ln_residual(a, w, bias=bias, scale=scale, shift=shift, residual=x, out_f32=x, out_bf16=xb)
cabi.ln_mod_residual(y, scale=scale, shift=shift, residual=x, out_f32=x, out_bf16=xb)
torch.cuda.synchronize()
