"""Timing probe for ab_window_attention on the three stage shapes of the 0.25-degree model (GPU box)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from aurora_b200 import cabi  # noqa: E402

WS = (2, 6, 12)
out = []
for name, res, heads in [("s1", (4, 180, 360), 8), ("s2", (4, 90, 180), 16), ("s3", (4, 45, 90), 32)]:
    l = res[0] * res[1] * res[2]
    d = heads * 64
    qkv = torch.randn(l, 3 * d, device="cuda").to(torch.bfloat16)
    pad = torch.randn(3 * d, device="cuda").to(torch.bfloat16)
    o = torch.empty(l, d, device="cuda", dtype=torch.bfloat16)
    for shifted in (False, True):
        ss = (1, 3, 6) if shifted else (0, 0, 0)
        for _ in range(3):
            cabi.window_attention(qkv, o, batch=1, res=res, window=WS, shift=ss, num_heads=heads, pad_qkv=pad)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            cabi.window_attention(qkv, o, batch=1, res=res, window=WS, shift=ss, num_heads=heads, pad_qkv=pad)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nw, nt, _ = cabi.window_geometry(res, WS, ss)
        fl = 4.0 * nw * heads * nt * nt * 64
        r = {"stage": name, "shifted": shifted, "ms": round(ms, 4), "gbs": round(8.0 * l * d / ms / 1e6, 1),
             "tflops": round(fl / ms / 1e9, 1)}
        print(json.dumps(r), flush=True)
        out.append(r)
