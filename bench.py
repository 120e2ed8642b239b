#!/usr/bin/env python
"""Benchmark of the Aurora forward hot path (BASELINE.json: forecast-steps/sec on the 0.25-degree
721x1440x13-level configuration, B = 1, history 2; 1.3 B-parameter `Aurora`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

* `--impl ours` (default): one "step" = one `Aurora.forward` through the sm_100a kernels.
    value : steps/s with the input Batch already resident in HBM (CUDA events, max over ranks)
    e2e   : steps/s through the public API starting from PINNED HOST tensors, H2D of every input field
            and D2H of the whole prediction inside the timed region
    roofline     : dominant kernel (tcgen05 GEMM): algorithmic FLOPs / CUDA-event time per launch vs the
                   measured cuBLAS bf16 peak of MEASURED_PEAKS.json; plus the attention and adaLN kernels
    cpu_baseline : the CPU oracle port timed on this box's host cores on a bounded sample (N = 1 only)
* `--impl reference`: the reference algorithm's CPU implementation (oracle port; the Python reference
  itself cannot travel to the GPU box) on the host cores, each step a bounded sample extrapolated by
  algorithmic FLOPs.
* N > 1: the forward pass does not need a collective for independent forecasts, so by default each rank runs
  its own replica of the workload ("weak" scaling, no data-path collective).  `--parallelism latshard` instead
  shards ONE forecast over the N GPUs by latitude band with an NCCL halo exchange per Swin block ("strong").

Synthetic data (ERA5-shaped, `loc + scale * N(0,1)` per variable / level) and random weights with the
reference's zero-initialised tensors re-drawn (otherwise every Swin block is an identity).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from datetime import datetime
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LEVELS13 = (50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000)

WORKLOADS = {
    # name: (model class, H, W, levels, description)
    "aurora-0.25deg-721x1440x13L": ("Aurora", 721, 1440, LEVELS13),
    "aurora-small-0.25deg-721x1440x13L": ("AuroraSmallPretrained", 721, 1440, LEVELS13),
    "aurora-small-17x32x4L": ("AuroraSmallPretrained", 17, 32, (100, 250, 500, 850)),
    "aurora-highres-0.1deg-1801x3600x13L": ("AuroraHighRes", 1801, 3600, LEVELS13),
    "aurora-airpollution-0.4deg-451x900x13L": ("AuroraAirPollution", 451, 900, LEVELS13),
    "aurora-wave-0.25deg-721x1440x13L": ("AuroraWave", 721, 1440, LEVELS13),
}
DEFAULT_WORKLOAD = "aurora-0.25deg-721x1440x13L"

# Algorithmic work of one forward step, 2 flops per MAC.
ALGO_TFLOP = {"aurora-0.25deg-721x1440x13L": 96.8, "aurora-highres-0.1deg-1801x3600x13L": 91.7,
              "aurora-airpollution-0.4deg-451x900x13L": 69.2, "aurora-wave-0.25deg-721x1440x13L": 96.8,
              "aurora-small-0.25deg-721x1440x13L": 12.6, "aurora-small-17x32x4L": 0.004}


def peaks() -> dict:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# synthetic inputs / weights
# ------------------------------------------------------------------------------------------------
def make_host_batch(cfg, h, w, levels, pinned: bool, seed: int = 0):
    """Physically scaled synthetic Batch on the host (optionally pinned)."""
    from aurora_b200 import Batch, Metadata
    from aurora_b200.stats import atmos_stats_of, surf_stats_of

    g = torch.Generator().manual_seed(seed)

    def alloc(shape):
        t = torch.empty(shape, dtype=torch.float32, pin_memory=pinned)
        t.normal_(generator=g)
        return t

    surf = {}
    surf_names = cfg.surf_vars
    wave = any(k.endswith("_density") for k in surf_names)
    if wave:  # AuroraWave takes the raw HRES-WAM names; absent wave components are NaN (here: a band of "land")
        surf_names = tuple(dict.fromkeys(k.removesuffix("_density").removesuffix("_sin").removesuffix("_cos")
                                         for k in surf_names))
    for k in surf_names:
        loc, sc = surf_stats_of(k)
        surf[k] = alloc((1, 2, h, w)).mul_(sc).add_(loc)
        if wave and k not in ("2t", "10u", "10v", "msl", "wind"):
            surf[k].abs_().clamp_(min=0.05)
            surf[k][..., :, : w // 4] = float("nan")
    static = {}
    for k in cfg.static_vars:
        loc, sc = surf_stats_of(k)
        static[k] = alloc((h, w)).mul_(sc).add_(loc)
        if k in ("wmb", "lat_mask"):
            static[k].copy_((static[k] > -1.0).float())
    atmos = {}
    for k in cfg.atmos_vars:
        locs, scs = atmos_stats_of(k, levels)
        t = alloc((1, 2, len(levels), h, w))
        t.mul_(torch.tensor(scs)[None, None, :, None, None]).add_(torch.tensor(locs)[None, None, :, None, None])
        atmos[k] = t
    meta = Metadata(lat=torch.linspace(90, -90, h), lon=torch.linspace(0, 360, w + 1)[:-1],
                    time=(datetime(2020, 6, 1, 12, 0),), atmos_levels=tuple(levels))
    return Batch(surf, static, atmos, meta)


def randomise_parameters_(model, seed: int = 0) -> None:
    """Random weights on the device: N(0, 0.02) matrices, the reference's zero-initialised adaLN
    modulation / LoRA-B / biases re-drawn so that every block contributes (timing is weight-independent)."""
    g = torch.Generator(device=next(model.parameters()).device).manual_seed(seed)
    for name, p in model.named_parameters():
        if name.endswith("norm.weight") or name.endswith(".2.weight") and p.dim() == 1 or name.endswith(".3.weight"):
            p.data.fill_(1.0)
        elif "ln_modulation" in name:
            p.data.normal_(0.0, 0.1, generator=g)
        elif p.dim() == 1:
            p.data.normal_(0.0, 0.02, generator=g)
        elif "token_embeds.weights" in name:
            p.data.uniform_(-0.1, 0.1, generator=g)
        else:
            p.data.normal_(0.0, 0.02, generator=g)
    for name, p in model.named_parameters():
        if p.dim() == 1 and ("surf_norm.weight" in name or name.endswith("norm.weight")):
            p.data.fill_(1.0)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = []
        for i, nm in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) — bounded sample
# ------------------------------------------------------------------------------------------------
_THREADS: dict = {}


def pick_threads(workload: str, cores: int) -> int:
    """Host threads for the CPU legs: the fastest of {cores, cores/2, cores/4} on a small slice of the sample
    (PyTorch's CPU kernels do not always scale to every hardware thread of a large box)."""
    if workload not in _THREADS:
        best = None
        for n in sorted({max(1, cores), max(1, cores // 2), max(1, cores // 4)}, reverse=True):
            t = cpu_baseline_sample(workload, n, lat_fraction=6, calibrate=True)["sample_seconds"]
            if best is None or t < best[0]:
                best = (t, n)
        _THREADS[workload] = best[1]
    return _THREADS[workload]


def cpu_baseline_sample(workload: str, cores: int, lat_fraction: int = 1, calibrate: bool = False) -> dict:
    """Time the CPU oracle on a bounded sample of the workload and extrapolate by algorithmic FLOPs.

    Sample: one Swin3D block of every U-Net stage (unshifted for stage 1, shifted for stages 2-3) at the
    workload's real widths and token grid, fp32, `cores` host threads; `lat_fraction` > 1 keeps only the first
    1 / lat_fraction of the latitude rows of every stage (whole window rows) to bound the sample further.
    Blocks are ~84 % of the step's FLOPs and cost the same ~1.7 TFLOP at every stage; the step time is the
    sample time scaled by (total step FLOPs / sample FLOPs)."""
    from oracle import aurora_oracle as O
    import aurora_b200 as ab

    cls, h, w, levels = WORKLOADS[workload]
    model_cfg = getattr(ab, cls)(_init="empty").config
    if not calibrate:
        cores = pick_threads(workload, cores)
    torch.set_num_threads(cores)
    p = model_cfg.patch_size
    res0 = (model_cfg.latent_levels, (h - h % p) // p, w // p)
    all_res, _ = O.encoder_specs(res0, len(model_cfg.encoder_depths))
    g = torch.Generator().manual_seed(0)
    sample_flops, t_total, parts = 0.0, 0.0, []
    c = torch.randn(1, model_cfg.embed_dim, generator=g)
    for i, res in enumerate(all_res):
        if lat_fraction > 1:
            wh = model_cfg.window_size[1]
            res = (res[0], max(wh, (res[1] // lat_fraction) // wh * wh), res[2])
        d = model_cfg.embed_dim * 2**i
        heads = model_cfg.encoder_num_heads[i]
        l = res[0] * res[1] * res[2]
        pre = "blk"
        sd = {
            f"{pre}.norm1.ln_modulation.1.weight": torch.randn(2 * d, model_cfg.embed_dim, generator=g) * 0.02,
            f"{pre}.norm1.ln_modulation.1.bias": torch.randn(2 * d, generator=g) * 0.1,
            f"{pre}.norm2.ln_modulation.1.weight": torch.randn(2 * d, model_cfg.embed_dim, generator=g) * 0.02,
            f"{pre}.norm2.ln_modulation.1.bias": torch.randn(2 * d, generator=g) * 0.1,
            f"{pre}.attn.qkv.weight": torch.randn(3 * d, d, generator=g) * 0.02,
            f"{pre}.attn.qkv.bias": torch.randn(3 * d, generator=g) * 0.02,
            f"{pre}.attn.proj.weight": torch.randn(d, d, generator=g) * 0.02,
            f"{pre}.attn.proj.bias": torch.randn(d, generator=g) * 0.02,
            f"{pre}.mlp.fc1.weight": torch.randn(4 * d, d, generator=g) * 0.02,
            f"{pre}.mlp.fc1.bias": torch.randn(4 * d, generator=g) * 0.02,
            f"{pre}.mlp.fc2.weight": torch.randn(d, 4 * d, generator=g) * 0.02,
            f"{pre}.mlp.fc2.bias": torch.randn(d, generator=g) * 0.02,
        }
        # (non-zero biases like any trained checkpoint: with zero biases the zero-padded tokens of shifted windows
        # produce denormal attention outputs, which slow the CPU GEMMs down several-fold)
        x = torch.randn(1, l, d, generator=g)
        cfg_nolora = type(model_cfg)(**{**model_cfg.__dict__, "use_lora": False})
        with torch.inference_mode():
            # a forward pass runs 12-20 blocks per stage: per-geometry setup (window map, shift mask, oneDNN primitive
            # creation) is paid once per stage there, so one untimed pass comes first and the warm pass is timed
            if not calibrate:
                O.swin_block(sd, pre, x, c, res, heads, i > 0, cfg_nolora, 0)
            t0 = time.perf_counter()
            O.swin_block(sd, pre, x, c, res, heads, i > 0, cfg_nolora, 0)
            dt = time.perf_counter() - t0
        nwin = 1
        ws, _ = O.W.adjust_windows(model_cfg.window_size, (0, 0, 0), res)
        pads = O.W.pad_lo_hi(res, ws)
        lp = 1
        for a in range(3):
            lp *= res[a] + pads[a][0] + pads[a][1]
        ntok = ws[0] * ws[1] * ws[2]
        flops = 2.0 * lp * d * 4 * d + 4.0 * (lp // ntok) * heads * ntok * ntok * 64 + 2.0 * l * d * 8 * d
        sample_flops += flops
        t_total += dt
        parts.append(f"stage{i + 1} {res} D={d}: {dt:.2f}s")
        del x, sd
    # The Perceiver encoder / decoder around the backbone are plain per-location MLP GEMMs (14 % of the
    # step's FLOPs at 0.25 degree) and run much closer to the CPU's GEMM peak than a Swin block does: sample them
    # separately (a slice of the decoder's Linear-GELU-Linear at its real widths) and extrapolate each part by its own
    # FLOPs.  Share of the blocks incl. patch merge / split: 82.86 of 96.8 TFLOP at cfg-Q; other
    # workloads use the same split.
    e = 2 * model_cfg.embed_dim
    hid = int(e * model_cfg.dec_mlp_ratio)
    rows = max(1024, (13 * res0[1] * res0[2]) // (16 * lat_fraction))
    msd = {"mlp.0.weight": torch.randn(hid, e, generator=g) * 0.02, "mlp.0.bias": torch.randn(hid, generator=g) * 0.02,
           "mlp.2.weight": torch.randn(e, hid, generator=g) * 0.02, "mlp.2.bias": torch.randn(e, generator=g) * 0.02}
    xm = torch.randn(1, rows, e, generator=g)
    with torch.inference_mode():
        if not calibrate:
            O._mlp(msd, "mlp", xm)
        t0 = time.perf_counter()
        O._mlp(msd, "mlp", xm)
        t_mlp = time.perf_counter() - t0
    mlp_flops = 4.0 * rows * e * hid
    parts.append(f"decoder MLP slice {rows}x{e}->{hid}->{e}: {t_mlp:.2f}s")
    step_flops = ALGO_TFLOP[workload] * 1e12
    block_share = 82.86 / 96.8
    est_step_s = (t_total * block_share * step_flops / sample_flops
                  + t_mlp * (1.0 - block_share) * step_flops / mlp_flops)
    t_total += t_mlp
    sample_flops += mlp_flops
    return {"value": 1.0 / est_step_s, "unit": "forecast-steps/s", "cores": cores, "kind": "port",
            "sample": f"one Swin3D block per U-Net stage at full width"
                      f"{'' if lat_fraction == 1 else f', first 1/{lat_fraction} of the latitude rows'} "
                      f"+ a slice of the Perceiver-decoder MLP ({'; '.join(parts)}; warm pass of two), fp32, "
                      f"{sample_flops / 1e12:.2f} of {step_flops / 1e12:.1f} TFLOP; step time extrapolated by FLOPs, "
                      f"blocks and Perceiver GEMMs separately, = {est_step_s:.1f} s",
            "sample_seconds": t_total}


def gemm_traffic(workload: str, launches: int, algo_bytes: float) -> dict:
    """`traffic` of the roofline object: DRAM bytes (read + write) per GEMM launch from the committed ncu pass
    over one step of this workload (profiles/gemm_traffic.json, written by tools/ncu_traffic.py from the ncu CSV;
    bench.py cannot run ncu itself), next to the compulsory bytes per launch counted from the launch arguments."""
    out = {"traffic": None, "algorithmic_bytes": algo_bytes / launches if launches else None}
    f = ROOT / "profiles" / "gemm_traffic.json"
    if f.exists():
        rec = json.loads(f.read_text()).get(workload)
        if rec:
            out["traffic"] = rec["dram_bytes_per_launch"]
            out["traffic_source"] = rec["source"]
    return out


# ------------------------------------------------------------------------------------------------
def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals, ms = [], []
    base = None
    # one full-size sample costs ~25 s on 128 cores; keep the whole run within a few minutes by shrinking the
    # latitude extent of the per-step sample when many steps are asked for
    n_samples = args.warmup + args.steps
    lat_fraction = 1 if n_samples <= 4 else 2 if n_samples <= 8 else 4
    for i in range(n_samples):
        base = cpu_baseline_sample(args.workload, cores, lat_fraction)
        if i >= args.warmup:
            vals.append(base["value"])
    v = statistics.mean(vals)
    cls, h, w, levels = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "forecast-steps/sec", "value": v, "unit": "forecast-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "model": cls, "grid": f"{h}x{w}", "levels": len(levels),
                   "note": "CPU oracle port of the reference algorithm; each step is a bounded sample "
                           "extrapolated by algorithmic FLOPs"},
        "cpu_baseline": {**base, "value": v},
        "e2e": {"value": v, "unit": "forecast-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--parallelism", default="replicas", choices=["replicas", "latshard"],
                    help="N > 1: independent forecasts per GPU (default) or ONE forecast sharded by latitude")
    ap.add_argument("--cuda-graph", action="store_true", help="model.use_cuda_graph = True (step replayed from a graph)")
    ap.add_argument("--rollout", type=int, default=0, metavar="N",
                    help="also time one N-step autoregressive rollout (BASELINE configs[2]: 40) and add a `rollout` object")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 1:
        args.warmup = 1

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import aurora_b200 as ab
    from aurora_b200 import cabi

    from aurora_b200 import dist as abd
    import torch.distributed as dist

    rank, world, local_rank = abd.env_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = abd.init_process_group("nccl", dev)

    cls, h, w, levels = WORKLOADS[args.workload]
    model = getattr(ab, cls)(_init="empty").to(dev).eval()
    randomise_parameters_(model, seed=rank)
    model.use_cuda_graph = bool(args.cuda_graph)
    cfg = model.config
    host_batch = make_host_batch(cfg, h, w, levels, pinned=True, seed=rank)
    dev_batch = host_batch.to(dev)
    h2d_bytes = sum(v.numel() * 4 for d in (host_batch.surf_vars, host_batch.static_vars, host_batch.atmos_vars)
                    for v in d.values())

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    latshard = distributed and args.parallelism == "latshard"
    if latshard:
        args.no_e2e = True  # the end-to-end leg is defined for whole forecasts per rank
        _fwd = model.forward
        model.forward = lambda b: _fwd(b, sharded=True)  # noqa: E731

    # ---- warm-up (packs weights, allocates workspace, caches encodings) ----
    for _ in range(args.warmup):
        pred = model.forward(dev_batch)
    d2h_bytes = sum(v.numel() * 4 for d in (pred.surf_vars, pred.atmos_vars) for v in d.values())
    host_out = {k: torch.empty(v.shape, dtype=torch.float32, pin_memory=True)
                for d in (pred.surf_vars, pred.atmos_vars) for k, v in d.items()}

    # ---- device-resident throughput ----
    barrier()
    launches0 = cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local_rank)   # samples through BOTH timed regions (device-resident and end-to-end)
    clk.__enter__()
    e0.record()
    for _ in range(args.steps):
        pred = model.forward(dev_batch)
    e1.record()
    barrier()
    launches = cabi.launch_count() - launches0
    ms = e0.elapsed_time(e1) / args.steps

    # ---- end to end through the public API from pinned host memory ----
    e2e_ms = None
    if not args.no_e2e:
        def e2e_step():
            p = model.forward(host_batch)     # the public call on HOST tensors: H2D of every field happens inside
                                              # (pinned source -> copy stream, overlaps the previous step's kernels)
            for grp in (p.surf_vars, p.atmos_vars):
                for k, v in grp.items():
                    host_out[k].copy_(v, non_blocking=True)   # D2H of the whole prediction
            return p
        e2e_step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1) / args.steps
    clk.__exit__()
    clocks = clk.summary()

    # ---- autoregressive rollout (rollout.py:14-49): state stays on the device, every prediction is yielded ----
    rollout_info = None
    if args.rollout > 0 and not latshard:
        def run_rollout(n):
            last = None
            for p in ab.rollout(model, host_batch, steps=n):   # initial H2D inside; the caller keeps only the last step
                last = p
            return last
        run_rollout(2)
        barrier()
        e0.record()
        run_rollout(args.rollout)
        e1.record()
        barrier()
        r_ms = e0.elapsed_time(e1)
        rollout_info = {"steps": args.rollout, "ms_total": r_ms, "ms_per_step": r_ms / args.rollout,
                        "value": world * 1000.0 * args.rollout / r_ms, "unit": "forecast-steps/s",
                        "note": "aurora_b200.rollout from a pinned host batch; history slide and predictions on the device"}

    # ---- per-kernel timing for the roofline (one instrumented step; CUDA events around each launch) ----
    cabi.PROFILE = {}
    model.use_cuda_graph = False
    model.forward(dev_batch)
    torch.cuda.synchronize()
    prof = cabi.PROFILE
    cabi.PROFILE = None
    pk = peaks()

    def agg(name):
        rows = prof.get(name, [])
        t = sum(a.elapsed_time(b) for a, b, _, _ in rows) / 1e3
        return len(rows), t, sum(r[2] for r in rows), sum(r[3] for r in rows)

    n_g, t_g, f_g, b_g = agg("gemm")
    n_a, t_a, f_a, b_a = agg("window_attention")
    n_l, t_l, _, b_l = agg("ln_mod_residual")
    peak_tf = pk["bf16_tflops_sustained"] or pk["bf16_tflops"]
    roofline = {
        "kernel": "gemm_bf16_tn_kernel (tcgen05)", "bound": "tensor",
        "achieved": f_g / t_g / 1e12 if t_g else None, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": (f_g / t_g / 1e12) / peak_tf if t_g else None,
        **gemm_traffic(args.workload, n_g, b_g),
        "peak_source": pk["source"] + " (sustained figure: kernel timed inside a long step)",
        "launches_per_step": n_g, "seconds_per_step": t_g, "share_of_step": t_g / (ms / 1e3),
        "others": {
            "window_attention": {
                "bound": "hbm", "launches_per_step": n_a, "seconds_per_step": t_a, "share_of_step": t_a / (ms / 1e3),
                "achieved_gbs": b_a / t_a / 1e9 if t_a else None, "hbm_peak_gbs": pk["hbm_gbs"],
                "frac_hbm": (b_a / t_a / 1e9) / pk["hbm_gbs"] if t_a else None,
                "achieved_tflops": f_a / t_a / 1e12 if t_a else None,
                "frac_tensor": (f_a / t_a / 1e12) / pk["bf16_tflops"] if t_a else None,
                # stand-alone roofline of this kernel: min(tensor peak, HBM peak x 72 FLOP/B)
                "standalone_roofline_tflops": min(pk["bf16_tflops"], pk["hbm_gbs"] * (f_a / b_a) / 1e3) if b_a else None,
                "frac_standalone_roofline": ((f_a / t_a / 1e12) / min(pk["bf16_tflops"], pk["hbm_gbs"] * (f_a / b_a) / 1e3)
                                             if t_a and b_a else None),
            },
            "ln_mod_residual": {
                "bound": "hbm", "launches_per_step": n_l, "seconds_per_step": t_l, "share_of_step": t_l / (ms / 1e3),
                "achieved_gbs": b_l / t_l / 1e9 if t_l else None,
                "frac_hbm": (b_l / t_l / 1e9) / pk["hbm_gbs"] if t_l else None,
            },
        },
    }

    # ---- max over ranks ----
    if distributed:
        mx = abd.max_over_ranks([ms, e2e_ms or 0.0], device=dev)
        ms, e2e_ms = mx[0], (mx[1] if e2e_ms is not None else None)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_sample(args.workload, os.cpu_count() or 1)
        algo_tflop = ALGO_TFLOP[args.workload]
        line = {
            "metric": "forecast-steps/sec", "value": (1 if latshard else world) * 1000.0 / ms, "unit": "forecast-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong" if latshard else "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": args.workload, "model_class": cls, "grid": f"{h}x{w}", "levels": len(levels),
                "batch_per_gpu": 1, "history": 2, "parameters_m": round(sum(p.numel() for p in model.parameters()) / 1e6, 1),
                "precision": "bf16 operands in the Swin backbone, fp16 operands in encoder/decoder, fp32 accumulate/"
                             "residual/LN/softmax",
                "parallelism": (f"one forecast latitude-sharded over {world} GPUs, NCCL halo exchange per block"
                                if latshard else f"replicas x{world}") if world > 1 else "single GPU",
                "l2_note": "inputs and activations are GBs per step (>> 126 MB L2); no explicit flush needed",
                "algorithmic_tflop_per_step": algo_tflop,
                "cuda_graph": bool(args.cuda_graph),
            },
            "model_tflops_achieved": algo_tflop / (ms / 1e3),
            "clocks": clocks,
            "gpu_launches": launches,
            "e2e": None if e2e_ms is None else {
                "value": world * 1000.0 / e2e_ms, "unit": "forecast-steps/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
            },
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        if rollout_info is not None:
            if distributed:
                rollout_info["note"] += " (rank 0's time)"
            line["rollout"] = rollout_info
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
