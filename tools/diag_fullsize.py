#!/usr/bin/env python
"""Where does the full-size parity error come from?  Runs OUR engine with stage taps and the unmodified reference
(oracle/_ref, fp32 and autocast) with forward hooks on the same modules, prints rel-mean-abs error per module.
    python tools/diag_fullsize.py [workload]"""
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
    model = getattr(ab, cls)(_init="empty", autocast=True).to(DEV).eval()
    bench.randomise_parameters_(model, seed=5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    host = bench.make_host_batch(model.config, h, w, levels, pinned=False, seed=5)
    dev_batch = host.to(DEV)
    eng = model._get_engine()
    eng.taps = {}
    model.forward(dev_batch)
    torch.cuda.synchronize()
    ours = {k: v[-1] for k, v in eng.taps.items()}
    eng.taps = None
    model._engine = None
    del model, eng
    gc.collect(); torch.cuda.empty_cache()
    rb = R.to_ref_batch(dev_batch)
    table = {}
    for tag, autocast in (("fp32", False), ("autocast", True)):
        with torch.device(DEV):
            rm = getattr(R.load(), cls)(autocast=autocast)
        rm.load_state_dict(sd, strict=True)
        rm.eval()
        got = {}
        hooks = []

        def add(name, mod, pick=lambda o: o):
            hooks.append(mod.register_forward_hook(lambda m, i, o, name=name, pick=pick: got.__setitem__(name, pick(o).detach().float()[0])))

        add("encoder", rm.encoder)
        add("backbone", rm.backbone)
        for kind, layers in (("encoder_layers", rm.backbone.encoder_layers), ("decoder_layers", rm.backbone.decoder_layers)):
            for i, layer in enumerate(layers):
                for j, blk in enumerate(layer.blocks):
                    add(f"backbone.{kind}.{i}.blocks.{j}", blk)
                if getattr(layer, "downsample", None) is not None:
                    add(f"backbone.{kind}.{i}.downsample", layer.downsample)
                if getattr(layer, "upsample", None) is not None:
                    add(f"backbone.{kind}.{i}.upsample", layer.upsample)
        with torch.inference_mode():
            rm.forward(rb)
        for hk in hooks:
            hk.remove()
        table[tag] = got
        del rm
        gc.collect(); torch.cuda.empty_cache()
    print(f"{'module':50s} {'ours vs fp32':>14s} {'ref-autocast vs fp32':>22s}  |ref| mean   max")
    for name, t in ours.items():
        key = name.replace("+skip", "")
        if key not in table["fp32"]:
            continue
        r32 = table["fp32"][key]
        ra = table["autocast"][key]
        if name.endswith("upsample+skip"):
            note = " (ours includes the additive skip)"
        else:
            note = ""
        if r32.shape != t.shape:
            print(f"{name:50s} shape {tuple(t.shape)} vs {tuple(r32.shape)}")
            continue
        print(f"{name:50s} {rel(t, r32):14.3e} {rel(ra, r32):22.3e}  {r32.abs().mean().item():9.3f} {r32.abs().max().item():9.1f}{note}")


if __name__ == "__main__":
    main()
