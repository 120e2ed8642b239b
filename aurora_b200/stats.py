"""Per-variable normalisation statistics (the constants of `aurora/normalisation.py:77-457`, stored in
``stats_table.json``) and the helpers that look them up."""

from __future__ import annotations

import json
from pathlib import Path
from typing import Optional

__all__ = ["locations", "scales", "level_to_str", "surf_stats_of", "atmos_stats_of"]

_table = json.loads((Path(__file__).with_name("stats_table.json")).read_text())
locations: dict[str, float] = _table["locations"]
scales: dict[str, float] = _table["scales"]


def level_to_str(level: float) -> str:
    """Canonical text form of a pressure level: ``850`` -> ``"850"``, ``12.5`` -> ``"12_5"``
    (`aurora/normalisation.py:17-31`)."""
    level = round(float(level), 3)
    if level % 1 == 0:
        level = int(level)
    return str(level).replace(".", "_")


def surf_stats_of(name: str, overrides: Optional[dict[str, tuple[float, float]]] = None) -> tuple[float, float]:
    """(location, scale) of a surface-level or static variable (`normalisation.py:34-49`)."""
    if overrides and name in overrides:
        loc, scale = overrides[name]
        return float(loc), float(scale)
    return float(locations[name]), float(scales[name])


def atmos_stats_of(name: str, levels) -> tuple[list[float], list[float]]:
    """Per-level (locations, scales) of an atmospheric variable (`normalisation.py:52-70`)."""
    locs = [float(locations[f"{name}_{level_to_str(lv)}"]) for lv in levels]
    scs = [float(scales[f"{name}_{level_to_str(lv)}"]) for lv in levels]
    return locs, scs
