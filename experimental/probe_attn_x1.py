#!/usr/bin/env python
"""Build experimental/libab_x1.so (attn_ptmem.cu + csrc/common.cu) and compare `ab_window_attention_x1` with the
shipped `ab_window_attention` on the three stage grids of the 0.25-degree model (needs a B200):

    timeout 300 python experimental/probe_attn_x1.py            # correctness + timing
    AB_X1_ONLY_BUILD=1 python experimental/probe_attn_x1.py     # just compile (works without a GPU)

Run it under `timeout`: the kernel has never executed (see the header of attn_ptmem.cu); a pipeline bug would show up
as a hang, which the kernels' own 4 s mbarrier watchdog turns into a trap."""
import ctypes as C
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
LIB = ROOT / "experimental" / "libab_x1.so"


def build() -> None:
    srcs = [ROOT / "experimental" / "attn_ptmem.cu", ROOT / "aurora_b200" / "csrc" / "common.cu"]
    if LIB.exists() and all(LIB.stat().st_mtime > s.stat().st_mtime for s in srcs):
        return
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
           "-Xcompiler", "-fPIC", "-shared", "-I", str(ROOT / "include"), "-I", str(ROOT / "aurora_b200" / "csrc"),
           *map(str, srcs), "-o", str(LIB), "-lcuda"]
    subprocess.run(cmd, check=True)


def main() -> None:
    build()
    if os.environ.get("AB_X1_ONLY_BUILD"):
        print("built", LIB)
        return
    import torch

    from aurora_b200 import cabi

    x1 = C.CDLL(str(LIB))
    x1.ab_last_error.restype = C.c_char_p
    ws = (2, 6, 12)
    for name, res, heads in [("s1", (4, 180, 360), 8), ("s2", (4, 90, 180), 16), ("s3", (4, 45, 90), 32)]:
        l, d = res[0] * res[1] * res[2], heads * 64
        g = torch.Generator(device="cuda").manual_seed(1)
        qkv = torch.randn(l, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
        pad = torch.randn(3 * d, device="cuda", generator=g).to(torch.bfloat16)
        for shifted in (False, True):
            ss = (1, 3, 6) if shifted else (0, 0, 0)
            ref = torch.empty(l, d, device="cuda", dtype=torch.bfloat16)
            out = torch.full((l, d), float("nan"), device="cuda", dtype=torch.bfloat16)
            cabi.window_attention(qkv, ref, batch=1, res=res, window=ws, shift=ss, num_heads=heads, pad_qkv=pad)
            a = cabi.AbWindowAttention()
            a.qkv, a.pad_qkv, a.out = qkv.data_ptr(), pad.data_ptr(), out.data_ptr()
            a.batch = 1
            a.res, a.window, a.shift = (C.c_int32 * 3)(*res), (C.c_int32 * 3)(*ws), (C.c_int32 * 3)(*ss)
            a.num_heads, a.head_dim, a.warped = heads, 64, 1
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

            def launch():
                rc = x1.ab_window_attention_x1(C.byref(a), stream)
                if rc != 0:
                    raise RuntimeError(x1.ab_last_error().decode())

            def timed(fn):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return round(e0.elapsed_time(e1) / 10, 4)

            rec = {"stage": name, "shifted": shifted}
            for variant in ("a", "b"):       # a = x1 (both P tiles in TMEM, 4 stages), b = x1b (P1 via smem, 3 stages)
                os.environ["AB_X1_VARIANT"] = variant
                out.fill_(float("nan"))
                launch()
                torch.cuda.synchronize()
                diff = (out.float() - ref.float()).abs()
                rec[f"x1{variant}"] = {"max_abs_diff": float(diff.max()), "nan": int(torch.isnan(out).sum()),
                                       "equal": bool(torch.equal(out, ref)), "ms": timed(launch)}
            rec["shipped_ms"] = timed(lambda: cabi.window_attention(
                qkv, ref, batch=1, res=res, window=ws, shift=ss, num_heads=heads, pad_qkv=pad))
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
