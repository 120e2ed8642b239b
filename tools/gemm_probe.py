"""Throughput probe for ab_gemm_bf16 on the shapes of Aurora's 0.25-degree forward (run on the GPU box)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from aurora_b200 import cabi  # noqa: E402

SHAPES = [
    ("qkv_s1", 259200, 1536, 512, 0), ("proj_s1", 259200, 512, 512, 0),
    ("fc1_s1", 259200, 2048, 512, 1), ("fc2_s1", 259200, 512, 2048, 0),
    ("qkv_s2", 64800, 3072, 1024, 0), ("fc1_s2", 64800, 4096, 1024, 1), ("fc2_s2", 64800, 1024, 4096, 0),
    ("qkv_s3", 16200, 6144, 2048, 0), ("fc1_s3", 16200, 8192, 2048, 1), ("fc2_s3", 16200, 2048, 8192, 0),
    ("proj_s2", 64800, 1024, 1024, 0), ("proj_s3", 16200, 2048, 2048, 0),
    ("dec_fc1", 842400, 2048, 1024, 1), ("dec_fc2", 842400, 1024, 2048, 0), ("dec_kv", 259200, 2048, 1024, 0),
    ("heads", 842400, 80, 1024, 0),
]
res = []
import os


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


for name, m, n, k, act in SHAPES:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") / k**0.5).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    r = {"name": name, "m": m, "n": n, "k": k, "act": act}
    # AB_GEMM_WIDE: unset = shipped heuristic, 0 = 256x256 pair tiles only, 2 = 512x256 tiles whenever K >= 1024
    for tag, env in (("", None), ("_pair", "0"), ("_wide", "2")):
        if env is None:
            os.environ.pop("AB_GEMM_WIDE", None)
        else:
            os.environ["AB_GEMM_WIDE"] = env
        ms = timed(lambda: cabi.gemm(a, w, bias=bias, out_bf16=out, act=act))
        r["ms" + tag] = round(ms, 4)
        r["tflops" + tag] = round(2.0 * m * n * k / ms / 1e9, 1)
    os.environ.pop("AB_GEMM_WIDE", None)
    ms_cublas = timed(lambda: torch.matmul(a, w.t()))  # cuBLAS on the same shape (no epilogue) for context
    r["cublas_ms"] = round(ms_cublas, 4)
    r["cublas_tflops"] = round(2.0 * m * n * k / ms_cublas / 1e9, 1)
    print(json.dumps(r), flush=True)
    res.append(r)
    del a, w, out
json.dump(res, open("gpurun_out/gemm_probe.json", "w"), indent=1)
