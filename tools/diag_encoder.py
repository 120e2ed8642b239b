#!/usr/bin/env python
"""Encoder-internal parity at full size: our engine's encoder buffers vs forward hooks on the unmodified reference's
encoder sub-modules (fp32).    python tools/diag_encoder.py [workload]"""
import gc
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import aurora_b200 as ab  # noqa: E402
import bench  # noqa: E402
from oracle import ref as R  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().mean() / b.abs().mean()).item()


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "aurora-small-0.25deg-721x1440x13L"
    cls, h, w, levels = bench.WORKLOADS[workload]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    global DEV
    dry = not torch.cuda.is_available()   # build container: exercise the reference-side hooks only
    if dry:
        DEV = "cpu"
    model = getattr(ab, cls)(_init="empty", autocast=True).to(DEV).eval()
    bench.randomise_parameters_(model, seed=5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    host = bench.make_host_batch(model.config, h, w, levels, pinned=False, seed=5)
    dev_batch = host.to(DEV)
    if dry:
        ours = None
    else:
        ours = run_ours(model, dev_batch, levels)
    d0 = model.config.embed_dim
    model._engine = None
    del model
    gc.collect()
    compare(ours, sd, cls, dev_batch, d0, h, w, levels)


def run_ours(model, dev_batch, levels):
    eng = model._get_engine()
    eng.taps = {}
    model.forward(dev_batch)
    torch.cuda.synchronize()
    buf = {nm: t for (nm, _, _), t in eng._buf.items()}
    return {"x": eng.taps["encoder"][-1], "xs0": buf["enc.xs0"].float().clone(), "xa": buf["enc.xa"].float().clone(),
            "lat1": buf["enc.lat1"].float().clone(), "posscale": eng._grid_cache[2].clone(), "lead": eng.lead_emb.clone(),
            "abs": next(iter(eng._abs_cache.values())).clone(),
            "lev_enc": eng._level_cache[tuple(levels)]["enc"].clone(),
            "A_surf": buf["enc.A_surf"].float().clone()}


def compare(ours, sd, cls, dev_batch, d0, h, w, levels):
    if DEV == "cuda":
        torch.cuda.empty_cache()
    with torch.device(DEV):
        rm = getattr(R.load(), cls)(autocast=False)
    rm.load_state_dict(sd, strict=True)
    rm.eval()
    got, ins = {}, {}
    enc = rm.encoder
    for name in ("surf_token_embeds", "atmos_token_embeds", "level_agg", "pos_embed", "scale_embed", "lead_time_embed",
                 "absolute_time_embed", "atmos_levels_embed", "surf_mlp", "surf_norm"):
        mod = getattr(enc, name)
        def hook(m, i, o, name=name):
            got[name] = o.detach().float()
            ins[name] = i[0].detach().float()

        mod.register_forward_hook(hook)

    def enc_hook(m, i, o):
        got["x"] = o.detach().float()[0]

    enc.register_forward_hook(enc_hook)
    with torch.inference_mode():
        rm.forward(R.to_ref_batch(dev_batch))
    sle = sd["encoder.surf_level_encoding"].float()
    print("shapes:", {k: tuple(v.shape) for k, v in got.items()})
    if ours is None:
        ours = {"x": got["x"], "xs0": got["surf_token_embeds"][0] + sle, "posscale": (got["pos_embed"] + got["scale_embed"]).reshape(-1, d0),
                "lead": got["lead_time_embed"].reshape(-1), "abs": got["absolute_time_embed"].reshape(-1),
                "lev_enc": got["atmos_levels_embed"].reshape(-1, d0),
                "xa": (got["atmos_token_embeds"] + got["atmos_levels_embed"].reshape(-1, 1, d0)).reshape(-1, d0)}
    L = ours["xs0"].shape[0]
    print(f"final encoder out      : all {rel(ours['x'], got['x']):.3e}  surf latent {rel(ours['x'][:L], got['x'][:L]):.3e}  "
          f"atmos latents {rel(ours['x'][L:], got['x'][L:]):.3e}")
    # per latitude band of the final output (token row = (c*H + h)*W + w)
    hp, wp = (h - h % 4) // rm.patch_size, w // rm.patch_size
    xo, xr = ours["x"].view(4, hp, wp, d0), got["x"].view(4, hp, wp, d0)
    for r0 in range(0, hp, max(1, hp // 6)):
        r1 = min(hp, r0 + max(1, hp // 6))
        print(f"   rows {r0:3d}-{r1:3d}: {rel(xo[:, r0:r1], xr[:, r0:r1]):.3e}")
    print(f"surf patch embed + enc : {rel(ours['xs0'], got['surf_token_embeds'][0] + sle):.3e}   |ref| {got['surf_token_embeds'].abs().mean():.3f}")
    at = got["atmos_token_embeds"]            # (B*C, L, D)
    lev = got["atmos_levels_embed"]           # (C, D) or (1, C, D)
    lev = lev.reshape(-1, d0)
    print(f"levels embed           : {rel(ours['lev_enc'], lev):.3e}")
    xa = ours["xa"].view(len(levels), L, d0)
    for ci in sorted({0, len(levels) // 2, len(levels) - 1}):
        print(f"atmos embed level {ci:2d}   : {rel(xa[ci], at[ci] + lev[ci]):.3e}   |ref| {at[ci].abs().mean():.3f}")
    ps = (got["pos_embed"] + got["scale_embed"]).reshape(-1, d0)
    print(f"pos+scale embed        : {rel(ours['posscale'], ps):.3e}   |ref| {ps.abs().mean():.3f}  pos alone |ref| {got['pos_embed'].abs().mean():.3f}")
    print(f"   pos-enc input  ours-vs-ref n/a; ref input range {ins['pos_embed'].min():.3f}..{ins['pos_embed'].max():.3f}")
    print(f"lead-time embed        : {rel(ours['lead'], got['lead_time_embed'].reshape(-1)):.3e}")
    print(f"abs-time embed         : {rel(ours['abs'].reshape(-1), got['absolute_time_embed'].reshape(-1)):.3e}")
    la = got["level_agg"]                     # (B*L, 3, D)
    print(f"level_agg out          : shape {tuple(la.shape)} |ref| {la.abs().mean():.3f}")
    ours_lat = xo[1:]                         # (3, hp, wp, D) includes pos/scale/time adds
    # reference: latents after aggregation, before pos/scale/time; ours lat1 is after LN1 only -> compare final minus adds
    add = (ps.view(hp, wp, d0)[None] + got["lead_time_embed"].reshape(1, 1, 1, d0) + got["absolute_time_embed"].reshape(1, 1, 1, d0))
    ours_agg = ours_lat - add
    ref_agg = la.view(hp, wp, 3, d0).permute(2, 0, 1, 3)
    print(f"aggregated latents     : {rel(ours_agg, ref_agg):.3e}")
    ours_surf = xo[0] - add[0]
    ref_surf = (ins["surf_norm"] * 0 + got["surf_norm"]).reshape(hp, wp, d0) + (got["surf_token_embeds"][0] + sle).view(hp, wp, d0)
    print(f"surface latent (pre add): {rel(ours_surf, ref_surf):.3e}")


if __name__ == "__main__":
    main()
