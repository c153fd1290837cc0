"""Image sharding across ranks for the test stage.

The reference shards images with the `DistributedSampler` Lightning injects into its bs=1 loader
(`pl_wrapper/sam2matcher_pl.py:231-239`): strided assignment, padded by repetition so every rank gets the same
count, and `collect_results_cpu` (`run_lightning.py:23-78`) re-interleaves the per-rank result lists with
`zip(*parts)` and truncates to the dataset length.  No collective runs while images are being scored.
"""
from __future__ import annotations

import math


def shard_indices(n_items: int, rank: int, world: int):
    """Indices of the images rank `rank` processes (DistributedSampler(shuffle=False, drop_last=False))."""
    if n_items == 0:
        return []
    per_rank = math.ceil(n_items / world)
    total = per_rank * world
    idx = list(range(n_items))
    while len(idx) < total:  # pad by repetition, like the sampler
        idx += idx[:total - len(idx)]
    return idx[rank:total:world]


def interleave(parts, n_items: int):
    """Inverse of `shard_indices` for the gathered per-rank result lists (run_lightning.py:69-75)."""
    ordered = []
    for group in zip(*parts):
        ordered.extend(group)
    return ordered[:n_items]


def collect_results(result_part, size=None, group=None):
    """`collect_results_cpu` (`run_lightning.py:23-78`) without the filesystem round trip: the per-rank result lists
    are gathered on rank 0 with one `gather_object` (NCCL/gloo), re-interleaved with `zip(*parts)` and truncated to
    `size` (the sampler may have padded).  Same contract: rank 0 gets the ordered list, every other rank `None`;
    without an initialised process group the part is returned unchanged.  With the fused RLE output a part holds a
    few KB per image, so the gather is a single small message."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return result_part
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(result_part, parts, dst=0, group=group)
    if rank != 0:
        return None
    ordered = []
    for res in zip(*parts):
        ordered.extend(list(res))
    if size is not None:
        ordered = ordered[:size]
    return ordered
