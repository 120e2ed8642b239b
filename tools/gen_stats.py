"""Regenerate aurora_b200/stats_table.json (per-variable normalisation locations and scales) from the
reference's tables (`aurora/normalisation.py:77-457`).  Run in the build container only:

    PYTHONPATH=/root/reference:tests/_shims python tools/gen_stats.py
"""
import json
from pathlib import Path

from aurora.normalisation import locations, scales

out = Path(__file__).resolve().parent.parent / "aurora_b200" / "stats_table.json"
out.write_text(json.dumps({"locations": locations, "scales": scales}, indent=0, sort_keys=True))
print(out, len(locations), len(scales))
