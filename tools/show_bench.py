"""Print the headline fields of a bench.py JSON line (helper of tools/gpu_round_check.sh)."""
import json
import sys

for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d = json.loads(ln)
        r = d.get("roofline") or {}
        o = r.get("others", {})
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("ms_per_step"), "launches",
              d.get("gpu_launches"), "graph", (d.get("details") or {}).get("cuda_graph"))
        print("  gemm TF/s", r.get("achieved"), "s/step", r.get("seconds_per_step"), "attn",
              (o.get("window_attention") or {}).get("seconds_per_step"), "ln", (o.get("ln_mod_residual") or {}).get("seconds_per_step"))
        for key in ("rollout", "replicas", "gpu_reference"):
            if d.get(key):
                print(" ", key, json.dumps(d[key])[:400])
