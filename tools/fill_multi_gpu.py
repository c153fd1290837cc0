#!/usr/bin/env python
"""Multi-GPU memory-bank fill check + timing (BASELINE config 3 shape: 80 classes x 30 shots, C=1024).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/fill_multi_gpu.py [--n-cls 80 --shots 30]

Each rank pools its strided share of the synthetic reference shots with `nttt_fill_pool_accumulate`, then
`MemoryBank.sync_fill()` runs the single NCCL all-reduce.  Rank 0 re-fills a single-process bank with ALL shots and
checks that the distributed result is bit-identical (single-writer slots: the sum only adds zeros).
"""
import argparse
import importlib
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-cls", type=int, default=80)
    ap.add_argument("--shots", type=int, default=30)
    ap.add_argument("--c", type=int, default=1024)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("no-time-to-train_b200")
    e = 37 * 37
    order = [(ci, li) for ci in range(args.n_cls) for li in range(args.shots)]  # L consecutive items per class
    gen = torch.Generator().manual_seed(7)

    def shot(ci, li):
        g = torch.Generator().manual_seed(1000 * ci + li)
        f = torch.randn(e, args.c, generator=g)
        m = torch.zeros(74, 74)
        y0, x0 = torch.randint(0, 30, (2,), generator=g).tolist()
        m[y0:y0 + 30, x0:x0 + 25] = 1.0
        m[y0, x0:x0 + 25] = 0.5
        return f, m

    bank = pkg.MemoryBank(dict(category_num=args.n_cls, length=args.shots, feat_shape=(e, args.c))).to(dev)
    mine = order[rank::world]
    data = [tuple(t.to(dev) for t in shot(ci, li)) for ci, li in mine]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for (ci, li), (f, m) in zip(mine, data):
        bank.fill(ci, f, m, (37, 37))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    bank.sync_fill()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    bank.postprocess()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    ok = None
    if rank == 0:
        ref = pkg.MemoryBank(dict(category_num=args.n_cls, length=args.shots, feat_shape=(e, args.c))).to(dev)
        was = dist.is_initialized()
        # single-process fill of everything, in the reference's arrival order (step-major, rank-minor)
        steps = (len(order) + world - 1) // world
        ref_fill = pkg.memory_bank.MemoryBank.fill
        seq = [order[s * world + r] for s in range(steps) for r in range(world) if s * world + r < len(order)]
        import unittest.mock as mock
        with mock.patch.object(dist, "is_initialized", return_value=False):
            for ci, li in seq:
                f, m = shot(ci, li)
                ref.fill(ci, f.to(dev), m.to(dev), (37, 37))
            ref.postprocess()
        ok = bool(torch.equal(ref.feats_ins_avg, bank.feats_ins_avg) and torch.equal(ref.feats_avg, bank.feats_avg)
                  and torch.equal(ref.fill_counts, bank.fill_counts) and torch.equal(ref.masks, bank.masks))
        print(json.dumps(dict(world=world, shots_total=len(order), shots_per_rank=len(mine),
                              fill_ms=1e3 * (t1 - t0), allreduce_ms=1e3 * (t2 - t1), postprocess_ms=1e3 * (t3 - t2),
                              allreduce_bytes=4 * (bank.feats_sum.numel() + bank.mask_sum.numel() + bank.masks.numel()),
                              bit_identical_to_single_process=ok)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
