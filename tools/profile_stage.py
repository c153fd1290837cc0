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
    ap.add_argument("--latency", action="store_true", help="time ONE image at a time: a CUDA-graph replay of the stage "
                    "in both launch modes (CUDA events, device idle before every replay) and the per-stage event profile")
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
    if args.latency:
        for low in (False, True):
            g = stage.graphed(args.n_masks, c, (1024, 1024), key=("lat", low), low_latency=low)
            g.lr_masks.copy_(images[0][0]); g.pred_ious.copy_(images[0][1]); g.tar_feat.copy_(images[0][2])
            g.capture()
            times = []
            for rep in range(24):
                img = images[rep % len(images)]
                g.lr_masks.copy_(img[0]); g.pred_ious.copy_(img[1]); g.tar_feat.copy_(img[2])
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record()
                torch.cuda.synchronize(dev)
                times.append(1e3 * e0.elapsed_time(e1))
            times = sorted(times[4:])
            print(f"graph replay, low_latency={low}: median {times[len(times) // 2]:.1f} us, min {times[0]:.1f} us per image")
        stage.profile(True)
        acc = {}
        for rep in range(12):
            p = stage.match_async(*images[rep % len(images)], (1024, 1024), slot=0, persistent_out=outs, low_latency=True)
            torch.cuda.synchronize(dev)
            if rep >= 2:
                for k, v in stage.profile_read().items():
                    acc[k] = acc.get(k, 0.0) + 1e3 * v / 10
        stage.profile(False)
        print("per-stage (eager, low-latency, includes host launch gaps): " + "  ".join(f"{k} {v:.1f}" for k, v in acc.items()))


if __name__ == "__main__":
    main()
