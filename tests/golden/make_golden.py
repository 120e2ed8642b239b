"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference, imported through the 4-symbol timm stand-in in tests/_shims) on the deterministic
inputs of tests/fixtures.py.  Run in the build container only (the GPU box has no /root/reference):

    PYTHONPATH=/root/reference:tests/_shims:. python tests/golden/make_golden.py

Outputs
    windows.npz        gather indices and mask group ids of roll->pad->window_partition_3d and
                       compute_3d_shifted_window_mask for the shape cases of SURVEY.md App. A.5
    model_<case>.npz   float32 copies of the reference's float64 forward outputs (+ encoder / backbone taps)
    keys.json          state_dict key -> shape for the reference presets (parameter-layout parity)
"""

import dataclasses
import hashlib
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

import aurora  # noqa: E402  (the reference)
from aurora.model import swin3d as ref_swin  # noqa: E402
from aurora.model.util import maybe_adjust_windows  # noqa: E402

from tests import fixtures as fx  # noqa: E402

WINDOW_CASES = [
    # (res, shifted)
    ((4, 45, 90), True), ((4, 45, 90), False), ((4, 4, 8), True), ((4, 4, 8), False),
    ((4, 15, 30), True), ((4, 15, 30), False), ((4, 38, 75), True), ((4, 38, 75), False),
    ((4, 12, 24), True), ((4, 12, 24), False), ((6, 13, 25), True), ((6, 13, 25), False),
    ((4, 8, 16), True), ((4, 2, 4), True), ((4, 1, 2), True), ((4, 90, 180), True), ((4, 75, 150), True),
]
# production shapes: only a checksum is stored
WINDOW_HASH_CASES = [((4, 180, 360), True), ((4, 180, 360), False), ((4, 150, 300), True), ((4, 90, 180), False)]
WS0 = (2, 6, 12)


def ref_window_maps(res, shifted, warped=True):
    c, h, w = res
    ss0 = tuple(s // 2 for s in WS0) if shifted else (0, 0, 0)
    ws, ss = maybe_adjust_windows(WS0, ss0, res)
    x = (torch.arange(c * h * w, dtype=torch.float64) + 1).view(1, c, h, w, 1)
    if not all(s == 0 for s in ss):
        x = torch.roll(x, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
    pad = ((-c) % ws[0], (-h) % ws[1], (-w) % ws[2])
    x = ref_swin.pad_3d(x, pad)
    idx = ref_swin.window_partition_3d(x, ws).reshape(-1, ws[0] * ws[1] * ws[2]).long() - 1
    if all(s == 0 for s in ss):
        groups = None
    else:
        ref_swin.compute_3d_shifted_window_mask.cache_clear()
        mask, img = ref_swin.compute_3d_shifted_window_mask(c, h, w, ws, ss, torch.device("cpu"), torch.float32, warped)
        groups = ref_swin.window_partition_3d(img, ws).reshape(-1, ws[0] * ws[1] * ws[2]).to(torch.uint8)
        # the mask is fully determined by the group ids
        same = groups[:, :, None] == groups[:, None, :]
        assert torch.equal(mask, torch.where(same, 0.0, -100.0))
    return idx.numpy(), None if groups is None else groups.numpy()


def gen_windows():
    out = {}
    for res, shifted in WINDOW_CASES:
        for warped in (True, False):
            idx, g = ref_window_maps(res, shifted, warped)
            tag = f"{res[0]}x{res[1]}x{res[2]}_{'s' if shifted else 'u'}_{'w' if warped else 'n'}"
            out[f"idx_{tag}"] = idx.astype(np.int32)
            if g is not None:
                out[f"grp_{tag}"] = g
    hashes = {}
    for res, shifted in WINDOW_HASH_CASES:
        idx, g = ref_window_maps(res, shifted, True)
        tag = f"{res[0]}x{res[1]}x{res[2]}_{'s' if shifted else 'u'}_w"
        hashes[f"idx_{tag}"] = hashlib.sha256(idx.astype(np.int32).tobytes()).hexdigest()
        if g is not None:
            hashes[f"grp_{tag}"] = hashlib.sha256(g.astype(np.uint8).tobytes()).hexdigest()
    np.savez_compressed(HERE / "windows.npz", **out)
    (HERE / "windows_hashes.json").write_text(json.dumps(hashes, indent=1))
    print("windows.npz", len(out), "arrays;", len(hashes), "hashes")


from tests.golden.cases import MODEL_CASES  # noqa: E402


def build_reference(cfg_name, cls_name, seed):
    cfg = fx.CONFIGS[cfg_name]
    cls = getattr(aurora, cls_name)
    model = cls(**fx.model_kwargs(cfg, cls_name))
    extra = fx.air_extra_specs(cfg) if cls_name == "AuroraAirPollution" else ()
    sd = fx.make_state_dict(cfg, seed=seed, extra=extra)
    model.load_state_dict(sd, strict=True)  # also proves key/shape parity of aurora_b200.spec
    model.eval()
    return cfg, model, sd


def to_ref_batch(b):
    return aurora.Batch(
        surf_vars=dict(b.surf_vars), static_vars=dict(b.static_vars), atmos_vars=dict(b.atmos_vars),
        metadata=aurora.Metadata(lat=b.metadata.lat, lon=b.metadata.lon, time=b.metadata.time,
                                 atmos_levels=b.metadata.atmos_levels, rollout_step=b.metadata.rollout_step),
    )


def gen_models(only=None):
    for name, (cfg_name, cls_name, h, w, levels, bsz, step, seed) in MODEL_CASES.items():
        if only and not any(name.startswith(o) for o in only):
            continue
        cfg, model, _ = build_reference(cfg_name, cls_name, seed)
        batch = fx.case_inputs(MODEL_CASES[name])[2]
        taps = {}
        h1 = model.encoder.register_forward_hook(lambda m, i, o: taps.__setitem__("encoder", o.detach()))
        h2 = model.backbone.register_forward_hook(lambda m, i, o: taps.__setitem__("backbone", o.detach()))
        model = model.double()
        with torch.inference_mode():
            pred = model.forward(to_ref_batch(batch))
        h1.remove(), h2.remove()
        out = {f"surf.{k}": v.float().numpy() for k, v in pred.surf_vars.items()}
        out.update({f"atmos.{k}": v.float().numpy() for k, v in pred.atmos_vars.items()})
        out["tap.encoder"] = taps["encoder"].float().numpy()
        out["tap.backbone"] = taps["backbone"].float().numpy()
        out["meta.rollout_step"] = np.array(pred.metadata.rollout_step)
        out["meta.time0"] = np.array(pred.metadata.time[0].timestamp())
        np.savez_compressed(HERE / f"model_{name}.npz", **out)
        print(name, "ok", sum(v.nbytes for v in out.values()) // 1024, "KiB")


def gen_rollout():
    cfg, model, _ = build_reference("tiny_lora", "Aurora", 7)
    batch = fx.make_batch(cfg, 33, 64, levels=fx.LEVELS4, b=1, seed=7)
    out = {}
    with torch.inference_mode():
        for i, pred in enumerate(aurora.rollout(model.double(), to_ref_batch(batch), steps=3)):
            for k, v in pred.surf_vars.items():
                out[f"step{i}.surf.{k}"] = v.float().numpy()
            for k, v in pred.atmos_vars.items():
                out[f"step{i}.atmos.{k}"] = v.float().numpy()
            out[f"step{i}.rollout_step"] = np.array(pred.metadata.rollout_step)
    np.savez_compressed(HERE / "rollout_tiny_lora.npz", **out)
    print("rollout ok")


def gen_keys():
    keys = {}
    for cls_name in ("Aurora", "AuroraPretrained", "AuroraSmallPretrained", "Aurora12hPretrained", "AuroraHighRes",
                     "AuroraAirPollution", "AuroraWave"):
        m = getattr(aurora, cls_name)()
        keys[cls_name] = {k: list(v.shape) for k, v in m.state_dict().items()}
    (HERE / "keys.json").write_text(json.dumps(keys))
    print("keys.json", {k: len(v) for k, v in keys.items()})


def gen_compat():
    """Published-layout checkpoints through the reference's `_adapt_checkpoint` of each model family."""
    from tests import compat_fixtures as cf

    out = {}
    for kind, cls_name in (("pretrained", "AuroraSmallPretrained"), ("air_pollution", "AuroraAirPollution"),
                           ("wave", "AuroraWave")):
        holder = types.SimpleNamespace(patch_size=cf.PATCH[kind])  # the adapters only read `self.patch_size`
        adapted = getattr(aurora, cls_name)._adapt_checkpoint(holder, cf.old_checkpoint(kind))
        out[kind] = cf.digest(adapted)
    (HERE / "compat.json").write_text(json.dumps(out, indent=0))
    print("compat.json", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    only = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--only=")]  # model case name prefixes
    what = [a for a in sys.argv[1:] if not a.startswith("--only=")] or ["windows", "models", "rollout", "keys", "compat"]
    torch.manual_seed(0)
    if "windows" in what:
        gen_windows()
    if "models" in what:
        gen_models(only)
    if "rollout" in what:
        gen_rollout()
    if "keys" in what:
        gen_keys()
    if "compat" in what:
        gen_compat()
