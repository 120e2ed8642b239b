"""Per-variable parity report of the CUDA path against the reference goldens (run on the GPU box)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, ".")
import aurora_b200 as ab  # noqa: E402
from tests import fixtures as fx  # noqa: E402
from tests.golden.cases import MODEL_CASES  # noqa: E402

out = {}
for name, (cfg_name, cls_name, h, w, levels, bsz, step, seed) in MODEL_CASES.items():
    cfg = fx.CONFIGS[cfg_name]
    model = getattr(ab, cls_name)(**fx.our_kwargs(cfg, cls_name))
    model.encoding_device = "cpu"  # goldens are CPU evaluations of the reference
    extra = fx.air_extra_specs(cfg) if cls_name == "AuroraAirPollution" else ()
    model.load_state_dict(fx.make_state_dict(cfg, seed=seed, extra=extra))
    model = model.to("cuda")
    batch = fx.make_batch(cfg, h, w, levels=levels, b=bsz, seed=seed, rollout_step=step)
    try:
        pred = model.forward(batch)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        out[name] = {"error": repr(e)}
        print(name, "ERROR", repr(e))
        continue
    gold = np.load(Path("tests/golden") / f"model_{name}.npz")
    errs = {}
    for grp, d in (("surf", pred.surf_vars), ("atmos", pred.atmos_vars)):
        for k, v in d.items():
            errs[f"{grp}.{k}"] = fx.rel_mean_abs(v.cpu(), torch.from_numpy(gold[f"{grp}.{k}"]))
    eng = model._get_engine()
    for tap, buf in (("encoder", "x0"),):
        pass
    out[name] = errs
    print(name, "worst", f"{max(errs.values()):.2e}", {k: f"{v:.1e}" for k, v in errs.items()})
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(out, open("gpurun_out/parity_report.json", "w"), indent=1)
