"""Launch each representative GEMM shape once (for ncu captures)."""
import sys
import torch
sys.path.insert(0, ".")
from aurora_b200 import cabi
for name, m, n, k, act in [("qkv_s1", 259200, 1536, 512, 0), ("fc1_s1", 259200, 2048, 512, 1), ("fc2_s1", 259200, 512, 2048, 0),
                           ("fc1_s2", 64800, 4096, 1024, 1), ("qkv_s2", 64800, 3072, 1024, 0), ("fc2_s3", 16200, 2048, 8192, 0)]:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") / k**0.5).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    cabi.gemm(a, w, bias=bias, out_bf16=out, act=act)
    torch.cuda.synchronize()
    if "--cublas" in sys.argv:
        torch.matmul(a, w.t(), out=out)  # the library kernel on the same operands, for side-by-side ncu captures
        torch.cuda.synchronize()
