"""Short eager run of the matching stage for ncu (no graphs: every kernel is a separate launch ncu can see).

    python tools/profile_stage.py [--images 3] [--n-masks 1024] [--n-classes 80] [--tune NAME=VALUE ...]

Config-2 shape by default (1024 masks, ViT-L features, 80 classes, 1024x1024 output); inputs come from the device
generator of `synth` (seed per image), prototypes from a synthetic clustered bank.
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=3)
    ap.add_argument("--n-masks", type=int, default=1024)
    ap.add_argument("--n-classes", type=int, default=80)
    ap.add_argument("--tune", action="append", default=[])
    args = ap.parse_args()
    pkg = importlib.import_module("no-time-to-train_b200")
    synth = pkg.synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    c = 1024
    centres = synth.cluster_centres(c)
    gen = torch.Generator().manual_seed(7)
    bank = centres[torch.arange(args.n_classes) % 5].unsqueeze(1) + (0.3 / c ** 0.5) * torch.randn(args.n_classes, 10, c, generator=gen)
    stage = pkg.MatchingStage(dev, pkg.StageConfig(nms_thr=0.5, num_out_instance=100, enc_hw=(37, 37)))
    stage.set_prototypes(bank)
    for item in args.tune:
        name, val = item.split("=")
        stage.tune(name, int(val))
    images = [synth.make_stage_inputs_device(args.n_masks, centres, dev, seed=1234 + i) for i in range(args.images)]
    torch.cuda.synchronize(dev)
    outs = (torch.zeros((100, 1024, 1024), dtype=torch.uint8, device=dev), torch.zeros((100, 4), dtype=torch.int32, device=dev))
    for rep in range(2):
        for img in images:
            p = stage.match_async(*img, (1024, 1024), slot=0, persistent_out=outs)
    torch.cuda.synchronize(dev)
    print("n_out", p.get()["counts"])


if __name__ == "__main__":
    main()
