"""aurora_b200 — Blackwell (sm_100a) implementation of Aurora's forward pass behind the reference's API
(`aurora/__init__.py:3-30`): ``Aurora*`` model classes, ``Batch``, ``Metadata`` and ``rollout``."""

from aurora_b200.batch import Batch, Metadata
from aurora_b200.model import (
    Aurora,
    Aurora12hPretrained,
    AuroraAirPollution,
    AuroraHighRes,
    AuroraPretrained,
    AuroraSmall,
    AuroraSmallPretrained,
    AuroraWave,
)
from aurora_b200.rollout import rollout

__all__ = [
    "Aurora", "AuroraPretrained", "AuroraSmallPretrained", "AuroraSmall", "Aurora12hPretrained", "AuroraHighRes",
    "AuroraAirPollution", "AuroraWave", "Batch", "Metadata", "rollout",
]
