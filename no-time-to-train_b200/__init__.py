"""no-time-to-train_b200 — B200-native (sm_100a) reference-matching stage of DogRog/no-time-to-train.

The directory name is not a Python identifier; import it with
`importlib.import_module("no-time-to-train_b200")` or through the `nttt_b200` alias module at the repo root.
"""
from . import _lib, synth  # noqa: F401
from .matching import MatchingStage, StageConfig  # noqa: F401
from .memory_bank import MemoryBank  # noqa: F401
from .model import Sam2MatchingBaselineNoAMG  # noqa: F401
from .results import encode_results  # noqa: F401
from .runner import MatcherRunner  # noqa: F401

__all__ = ["MatchingStage", "StageConfig", "MemoryBank", "Sam2MatchingBaselineNoAMG", "encode_results", "MatcherRunner", "synth"]
