"""End-to-end parity at BASELINE.json's FULL configurations: the whole `forward` of the production models on the
production grids — 48 Swin blocks at D = 512 ... 2048, the fp16 encoder / decoder on physically scaled 13-level inputs,
the decoder's 842 400-row GEMMs — against the UNMODIFIED reference (`oracle/_ref`, `oracle/build_ref.py`) running in
fp32 (no TF32) on the same GPU with the same parameters and the same Batch.  This is the reference's own acceptance
test (`tests/test_model.py:27-86`) moved to the benchmarked configurations.

Tolerance, per variable, rel-mean-abs error mean|out - ref| / mean|ref|:
    err(ours vs reference fp32)  <=  max( REF_TOL[var],  1.5 x err(reference autocast=True vs reference fp32) )
`REF_TOL` is the reference's own per-variable budget (1e-4 for 2t / msl / t / z, 5e-3 for the rest, `tests/fixtures.py`).
The second term states the claim precisely: our bf16 path must sit inside the spread of the reference's OWN bf16 recipe
(`Aurora(autocast=True)`, aurora.py:327-343) measured on the very same inputs — at 48 blocks that recipe itself can leave
the tight 1e-4 budget, which was written for the 20-block small model in float64.  Both errors are printed.

A stage-level tap is compared too: the backbone output (what `Perceiver3DDecoder` consumes) of the reference, captured
with a forward hook, against the engine's decoder input buffer."""

import gc

import pytest
import torch

from tests import fixtures as fx

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = {
    # name: (bench workload, reference / our class)
    "aurora_1p3b_0p25deg": "aurora-0.25deg-721x1440x13L",
    "aurora_small_0p25deg": "aurora-small-0.25deg-721x1440x13L",
    "aurora_airpollution_0p4deg": "aurora-airpollution-0.4deg-451x900x13L",
    "aurora_highres_0p1deg": "aurora-highres-0.1deg-1801x3600x13L",
}


def _rel(out: torch.Tensor, ref: torch.Tensor) -> float:
    out, ref = out.double(), ref.double()
    return ((out - ref).abs().mean() / ref.abs().mean()).item()


def _free():
    gc.collect()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", sorted(CASES))
def test_full_size_forward_matches_the_unmodified_reference(name):
    import aurora_b200 as ab
    import bench
    from oracle import ref as R

    if not R.available():
        pytest.skip("oracle/_ref was not built (run oracle/build_ref.py in the build container)")
    workload = CASES[name]
    cls, h, w, levels = bench.WORKLOADS[workload]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False   # conv3d patch embedding of the reference in true fp32

    # ---- ours -------------------------------------------------------------------------------------------
    model = getattr(ab, cls)(_init="empty", autocast=True).to(DEV).eval()
    bench.randomise_parameters_(model, seed=5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    host = bench.make_host_batch(model.config, h, w, levels, pinned=False, seed=5)
    dev_batch = host.to(DEV)
    pred = model.forward(dev_batch)
    torch.cuda.synchronize()
    ours = {("surf", k): v.clone() for k, v in pred.surf_vars.items()}
    ours.update({("atmos", k): v.clone() for k, v in pred.atmos_vars.items()})
    eng = model._engine
    tap_ours = next(t for (nm, _, _), t in eng._buf.items() if nm == "bb.concat").float().clone()
    meta = pred.metadata
    del pred, eng
    model._engine = None
    del model
    _free()

    # ---- the unmodified reference, fp32 and its own autocast recipe, same GPU ----------------------------------
    rb = R.to_ref_batch(dev_batch)
    taps = {}
    errs = {}
    for tag, autocast in (("fp32", False), ("autocast", True)):
        with torch.device(DEV):
            rmodel = getattr(R.load(), cls)(autocast=autocast)
        rmodel.load_state_dict(sd, strict=True)
        rmodel = rmodel.eval()
        hook = rmodel.backbone.register_forward_hook(lambda m, i, o, tag=tag: taps.__setitem__(tag, o.detach().float()[0]))
        with torch.inference_mode():
            rp = rmodel.forward(rb)
        hook.remove()
        torch.cuda.synchronize()
        outs = {("surf", k): v for k, v in rp.surf_vars.items()}
        outs.update({("atmos", k): v for k, v in rp.atmos_vars.items()})
        if tag == "fp32":
            ref_out = {k: v.clone() for k, v in outs.items()}
            assert rp.metadata.time == meta.time and rp.metadata.rollout_step == meta.rollout_step
            assert rp.metadata.atmos_levels == meta.atmos_levels
        else:
            errs["ref_autocast"] = {k: _rel(v, ref_out[k]) for k, v in outs.items()}
        del rp, outs, rmodel
        _free()

    assert sorted(ours) == sorted(ref_out)
    failures = []
    print(f"\n[full-size parity] {name} ({cls}, {h}x{w}x{len(levels)}L)")
    tap_err = _rel(tap_ours, taps["fp32"])
    tap_ref = _rel(taps["autocast"], taps["fp32"])
    print(f"  backbone output tap: ours {tap_err:.3e}   reference autocast {tap_ref:.3e}")
    for key in sorted(ours):
        o, r = ours[key], ref_out[key]
        assert o.shape == r.shape and torch.isfinite(o).all(), key
        e = _rel(o, r)
        e_ref = errs["ref_autocast"][key]
        bound = max(fx.tol_for(key[1]), 1.5 * e_ref)
        print(f"  {key[0]:5s} {key[1]:6s} ours {e:.3e}   reference autocast {e_ref:.3e}   bound {bound:.1e}"
              f"   {'(within the tight per-variable budget)' if e <= fx.tol_for(key[1]) else ''}")
        if not e <= bound:
            failures.append((key, e, bound))
    assert not failures, failures
    assert tap_err <= max(5e-3, 1.5 * tap_ref), (tap_err, tap_ref)
