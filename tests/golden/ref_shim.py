"""Import shim that lets the REAL reference hot-path modules import in the authoring container.

Only used by ``tests/golden/make_golden.py`` (and optional local cross-checks); it never travels to the
GPU box in a way that matters because ``/root/reference`` does not exist there.  The reference imports
hydra / omegaconf / matplotlib / pycocotools / tidecv at package-import time
(``sam2/__init__.py:7-9``, ``no_time_to_train/models/matching_baseline_utils.py:4``,
``no_time_to_train/dataset/visualization.py:8-9``); none of them is on the matching path, so each is
replaced by an empty stub module whose attributes resolve to inert placeholders.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NTTT_REFERENCE_ROOT", "/root/reference")


class _Anything:
    """Placeholder object: callable, attribute-able, iterable-empty."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


_STUBS = [
    "hydra", "hydra.utils", "hydra.core", "hydra.core.global_hydra",
    "omegaconf",
    "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
    "matplotlib.font_manager", "matplotlib.cm",
    "pycocotools", "pycocotools.coco", "pycocotools.cocoeval", "pycocotools.mask",
    "tidecv", "tidecv.datasets",
    "huggingface_hub",
]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "no_time_to_train"))


def install() -> None:
    """Register the stubs and put the reference on sys.path (idempotent)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = _StubModule(name)
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        mod.__path__ = []  # behave like a package so that submodule imports resolve
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Return the real reference symbols used on the matching path."""
    install()
    from no_time_to_train.models import matching_baseline_utils as mbu
    from no_time_to_train.models import Sam2MatchingBaseline_noAMG as model_mod
    from no_time_to_train.models import model_utils
    from sam2.utils import amg
    return types.SimpleNamespace(
        MemoryBank=mbu.MemoryBank,
        compute_sim_global_avg=mbu.compute_sim_global_avg,
        compute_sim_global_avg_with_neg=mbu.compute_sim_global_avg_with_neg,
        compute_semantic_ios=mbu.compute_semantic_ios,
        Model=model_mod.Sam2MatchingBaselineNoAMG,
        concat_all_gather=model_utils.concat_all_gather,
        batched_mask_to_box=amg.batched_mask_to_box,
        calculate_stability_score=amg.calculate_stability_score,
    )
