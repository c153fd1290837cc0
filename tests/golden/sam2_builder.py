"""Random-init SAM-2 + DINOv2 builders for tests that run the REAL reference model in the authoring container.

The reference builds SAM-2 through Hydra (`sam2/build_sam.py:44-78`: `compose` + `instantiate`) and DINOv2 through
`AutoModel.from_pretrained` (`no_time_to_train/models/model_utils.py:35-72`); neither hydra nor the network is
available here.  `build_predictor` is the ~15-line `_target_` instantiator SURVEY.md §8c describes: it loads the
reference's own `sam2_configs/<cfg>.yaml`, applies the overrides `build_sam2_video_predictor` passes, and instantiates
the reference's own classes with their default (random) initialisation.  Only reachable when /root/reference exists.
"""
import importlib
import os

import torch
import yaml

import ref_shim


def _coerce(v):
    if isinstance(v, str):
        try:
            return float(v)  # PyYAML 1.1 reads `1e-6` (no dot) as a string
        except ValueError:
            return v
    return v


def _instantiate(node):
    if isinstance(node, dict):
        kwargs = {k: _instantiate(v) for k, v in node.items() if k != "_target_"}
        if "_target_" in node:
            mod, _, name = node["_target_"].rpartition(".")
            return getattr(importlib.import_module(mod), name)(**kwargs)
        return kwargs
    if isinstance(node, list):
        return [_instantiate(v) for v in node]
    return _coerce(node)


def build_predictor(cfg_name="sam2_hiera_t.yaml", seed=0):
    ref_shim.install()
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "sam2_configs", cfg_name)) as f:
        cfg = yaml.safe_load(f)["model"]
    cfg["_target_"] = "sam2.sam2_video_predictor.SAM2VideoPredictor"
    cfg["sam_mask_decoder_extra_args"] = dict(dynamic_multimask_via_stability=True,
                                              dynamic_multimask_stability_delta=0.05,
                                              dynamic_multimask_stability_thresh=0.98)
    cfg["binarize_mask_from_pts_for_mem_enc"] = True
    cfg["fill_hole_area"] = 8
    torch.manual_seed(seed)
    return _instantiate(cfg).eval()


def build_dinov2(hidden=384, layers=2, heads=6, seed=1):
    from transformers import Dinov2Config, Dinov2Model
    torch.manual_seed(seed)
    return Dinov2Model(Dinov2Config(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                                    image_size=518, patch_size=14)).eval()
