"""Marginal cost of each stage of nttt_match_image with images in flight (one process, one synthetic pool).

    python tools/ablate.py [--batch 16] [--streams 16] [--steps 8] > gpurun_out/ablate.jsonl

For k = 1..n_stages the pipeline is cut after its k-th stage (NTTT_STOP_AFTER, read at nttt_ctx creation: a fresh ctx
is created per point) and the throughput of the truncated stage is measured exactly like bench.py's `value`
(CUDA-graph replay, `--streams` images in flight, inputs larger than L2).  The difference between consecutive points
is what a stage costs when the GPU is shared with the other images' kernels — not its isolated duration.
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["lowres_pack", "project_masks", "pool_gemm", "normalize_rows", "sim_top1", "box_nms", "upsample_pack",
          "mask_ios", "decay_rank", "unpack"]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--streams", type=int, default=16)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--n-masks", type=int, default=1024)
    ap.add_argument("--stops", default="", help="comma-separated stage numbers to measure (default: all, then 0 = whole)")
    args = ap.parse_args()
    pkg = importlib.import_module("no-time-to-train_b200")
    ops = importlib.import_module("no-time-to-train_b200.ops")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, S = args.batch, min(args.streams, args.batch)
    pool = [pkg.synth.make_stage_inputs(args.n_masks, 1024, 80, 10, (1024, 1024), seed=1234 + i, clustered=True)
            for i in range(B)]
    resident = [(p.lr_masks.to(dev), p.pred_ious.to(dev), p.tar_feat.to(dev)) for p in pool]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    prev = 0.0
    stops = [int(x) for x in args.stops.split(",")] if args.stops else list(range(1, len(STAGES) + 1)) + [0]
    if not pkg._lib.load().nttt_build_is_ablation():
        raise SystemExit("tools/ablate.py needs the ablation build: NTTT_BUILD_ABLATE=1 python no-time-to-train_b200/build.py")
    for stop in stops:
        os.environ["NTTT_STOP_AFTER"] = str(stop)
        ops._ctx_by_device.pop(dev.index, None)  # a fresh ctx reads the variable (the old one is leaked: tool only)
        stage = pkg.MatchingStage(dev, pkg.StageConfig(nms_thr=0.5, num_out_instance=100, enc_hw=(37, 37)))
        stage.set_prototypes(pool[0].feats_ins_avg)
        graphs = []
        for i in range(B):
            g = stage.graphed(args.n_masks, 1024, (1024, 1024), key=("ablate", stop, i))
            g.lr_masks, g.pred_ious, g.tar_feat = resident[i]
            graphs.append(g.capture())

        def step():
            for i in range(B):
                with torch.cuda.stream(streams[i % S]):
                    graphs[i].replay()

        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        cur = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for s in streams:
            s.wait_event(e0)
        for _ in range(args.steps):
            step()
        for s in streams:
            cur.wait_stream(s)
        e1.record(cur)
        torch.cuda.synchronize(dev)
        us = 1e3 * e0.elapsed_time(e1) / (args.steps * B)
        name = STAGES[stop - 1] if stop else "all (incl. rle if enabled)"
        print(json.dumps(dict(stop_after=stop, last_stage=name, us_per_image=round(us, 2),
                              marginal_us=round(us - prev, 2), streams=S, batch=B)), flush=True)
        prev = us
        del graphs, stage
        torch.cuda.synchronize(dev)


if __name__ == "__main__":
    main()
