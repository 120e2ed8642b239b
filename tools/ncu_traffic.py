#!/usr/bin/env python
"""Summarise an ncu metrics CSV of `bench.py` (long format: one row per launch x metric):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\
        --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
    python tools/ncu_traffic.py gpurun_out/launches.csv --tag r01m [--workload NAME]

writes  profiles/<tag>_launch_list.md   per-kernel launches / time share / DRAM bytes (all steps captured, equal shares)
        profiles/gemm_traffic.json      DRAM bytes per GEMM launch, read by bench.py for `roofline.traffic`
The numbers under ncu are cold-cache and serialised: shares and bytes are meaningful, absolute times are not."""

import argparse
import csv
import gzip
import json
import re
import shutil
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)            # drop the parameter list
    name = re.sub(r"^void ", "", name)
    return name[:110]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--workload", default="aurora-0.25deg-721x1440x13L")
    args = ap.parse_args()
    opener = gzip.open if args.csv.endswith(".gz") else open
    with opener(args.csv, "rt", newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    per_launch: dict = defaultdict(dict)
    for r in rows:
        val = float(r["Metric Value"].replace(",", "") or 0)
        unit = r["Metric Unit"]
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per_launch[int(r["ID"])]["name"] = r["Kernel Name"]
        per_launch[int(r["ID"])][r["Metric Name"]] = val * scale
    agg: dict = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    mine = set()
    for rec in per_launch.values():
        if "ab::" in rec["name"]:  # also kernels of nested namespaces whose printed name drops the `ab::` prefix
            mine.add(short(rec["name"]))
        a = agg[short(rec["name"])]
        a[0] += 1
        a[1] += rec.get("gpu__time_duration.sum", 0.0)
        a[2] += rec.get("dram__bytes_read.sum", 0.0)
        a[3] += rec.get("dram__bytes_write.sum", 0.0)
    ours = {k: v for k, v in agg.items() if k in mine}
    total_ns = sum(v[1] for v in ours.values())
    out = [f"# {args.tag} — ncu launch list of `bench.py --steps 1 --warmup 1` ({args.workload})",
           f"kernels of libaurora_b200.so only; {sum(v[0] for v in ours.values())} launches over the captured steps "
           f"(warm-up + timed + instrumented step), sum of gpu__time_duration = {total_ns / 1e6:.2f} ms "
           "(ncu: cold cache, serialised — compare shares, not absolutes)", "",
           "| kernel | launches | ms | share | DRAM read GB | DRAM write GB | DRAM bytes / launch (MB) |",
           "|---|---:|---:|---:|---:|---:|---:|"]
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / total_ns:.1f}% | {v[2] / 1e9:.2f} | "
                   f"{v[3] / 1e9:.2f} | {(v[2] + v[3]) / v[0] / 1e6:.1f} |")
    other = {k: v for k, v in agg.items() if k not in mine}
    out += ["", f"other kernels (PyTorch fills / copies): {sum(v[0] for v in other.values())} launches, "
                f"{sum(v[1] for v in other.values()) / 1e6:.2f} ms"]
    (ROOT / "profiles" / f"{args.tag}_launch_list.md").write_text("\n".join(out) + "\n")
    gem = [v for k, v in ours.items() if "gemm" in k]
    n = sum(v[0] for v in gem)
    if n:
        f = ROOT / "profiles" / "gemm_traffic.json"
        rec = json.loads(f.read_text()) if f.exists() else {}
        rec[args.workload] = {
            "dram_bytes_per_launch": sum(v[2] + v[3] for v in gem) / n, "launches": n,
            "dram_read_bytes": sum(v[2] for v in gem), "dram_write_bytes": sum(v[3] for v in gem),
            "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over all {n} GEMM launches of the captured steps "
                      f"(profiles/{args.tag}_launch_list.md)",
        }
        f.write_text(json.dumps(rec, indent=1) + "\n")
    dst = ROOT / "profiles" / f"{args.tag}_launches.csv.gz"
    if not args.csv.endswith(".gz"):
        with open(args.csv, "rb") as fi, gzip.open(dst, "wb") as fo:
            shutil.copyfileobj(fi, fo)
    print("\n".join(out[:12]))


if __name__ == "__main__":
    sys.exit(main())
