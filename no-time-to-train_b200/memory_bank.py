"""Memory bank of the matching model, B200-native.

Mirror of `MemoryBank` (`no_time_to_train/models/matching_baseline_utils.py:538-656`): same constructor, same
`postprocess()` entry, same state-dict names for everything the test path reads (`fill_counts`, `masks`,
`feats_avg`, `feats_ins_avg`, `postprocessed`) so Lightning checkpoints interchange (`load_state_dict(strict=False)`,
`pl_wrapper/sam2matcher_pl.py:140-142`).

What differs, on purpose:
  * the raw `feats [n_cls, L, E, C]` buffer (4.5 GB at 80x10 ViT-L, 67 GB at LVIS scale) is never stored.  Each
    reference shot is reduced on arrival by `nttt_fill_pool_accumulate` to `feats_sum[c,l,:] = sum_e m[e] f[e,:]`
    and `mask_sum[c,l] = sum_e m[e]`; `postprocess` is then a few KB of arithmetic (`nttt_fill_finalize`).
  * multi-GPU fill: the reference all-gathers the raw 5.6 MB feature map of every shot on every step
    (`Sam2MatchingBaseline_noAMG.py:471-474`).  Here each rank pools its own shots locally; `sync_fill()` does one
    tiny all_gather of the (step, class) log to reproduce the reference's arrival-order slot assignment, scatters
    the local pooled sums into their slots and issues ONE `all_reduce(SUM)` (NCCL over NVLink).  Every slot has a
    single writer, so the reduction only adds zeros and the result is bit-identical for any world size.
  * `feats_covariances`, `feats_centers`, `ins_sim_avg`, `pca_*` (never read by fill/test, only by the
    out-of-scope `vis_memory`) are not computed.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops


class MemoryBank(nn.Module):
    def __init__(self, config, kmeans_k=None, n_pca_components=None):
        super().__init__()
        self.n_classes = config.get("category_num")
        self.length = config.get("length")
        self.feat_shape = config.get("feat_shape")
        self.kmeans_k = kmeans_k
        self.n_pca_components = n_pca_components
        assert len(self.feat_shape) == 2
        _mem_n, _mem_c = self.feat_shape
        self.register_buffer("fill_counts", torch.zeros((self.n_classes,), dtype=torch.long))
        self.register_buffer("masks", torch.zeros((self.n_classes, self.length, _mem_n)))
        self.register_buffer("feats_avg", torch.zeros((self.n_classes, _mem_c)))
        self.register_buffer("feats_ins_avg", torch.zeros((self.n_classes, self.length, _mem_c)))
        self.register_buffer("postprocessed", torch.zeros((1,), dtype=torch.bool))
        # compact replacement of the raw `feats` buffer
        self.register_buffer("feats_sum", torch.zeros((self.n_classes, self.length, _mem_c)))
        self.register_buffer("mask_sum", torch.zeros((self.n_classes, self.length)))
        self.ready = False
        self._host_counts = None  # host mirror of fill_counts (avoids a device sync per shot)
        self._staged = []         # distributed fill: (class, pooled_sum [C], mask_sum [1], mask [E]) per local step
        self.register_state_dict_pre_hook(lambda module, prefix, keep_vars: module.sync_fill())

    # ------------------------------------------------------------------------------------------------
    def _counts(self):
        if self._host_counts is None:
            self._host_counts = self.fill_counts.tolist()
        return self._host_counts

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._host_counts = None

    def fill(self, cat_ind: int, feat: torch.Tensor, soft_mask: torch.Tensor, enc_hw) -> None:
        """One reference shot (`forward_fill_memory`, `Sam2MatchingBaseline_noAMG.py:465-485`).

        feat [E, C] f32 encoder output, soft_mask [S, S] f32 in [0,1] (nearest-resized to enc_hw inside the
        kernel, like `F.interpolate(mode="nearest")` at :465-469)."""
        feat = feat.reshape(-1, feat.shape[-1]).contiguous()
        soft_mask = soft_mask.reshape(soft_mask.shape[-2], soft_mask.shape[-1]).to(feat.device, torch.float32).contiguous()
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if distributed:
            assert (self.n_classes * self.length) % dist.get_world_size() == 0  # :479-480
            pooled = torch.zeros((feat.shape[-1],), dtype=torch.float32, device=feat.device)
            wsum = torch.zeros((1,), dtype=torch.float32, device=feat.device)
            m = ops.fill_pool_accumulate(feat, soft_mask, enc_hw, pooled, wsum, want_mask=True)
            self._staged.append((int(cat_ind), pooled, wsum, m))
            return
        counts = self._counts()
        slot = counts[cat_ind]
        if slot >= self.length:
            raise IndexError(f"memory bank slot overflow for class {cat_ind}")  # the reference raises IndexError too
        m = ops.fill_pool_accumulate(feat, soft_mask, enc_hw, self.feats_sum[cat_ind, slot], self.mask_sum[cat_ind, slot:slot + 1],
                                     want_mask=True)
        self.masks[cat_ind, slot] += m
        counts[cat_ind] += 1
        self.fill_counts[cat_ind] += 1

    def sync_fill(self) -> None:
        """Resolve the staged distributed fill: slot assignment in the reference's arrival order, then one
        all_reduce.  No-op when nothing is staged on any rank (checked collectively only if dist is up)."""
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if not distributed:
            return
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = self.feats_sum.device
        n_local = torch.tensor([len(self._staged)], dtype=torch.long, device=dev)
        n_all = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(n_all, n_local)
        n_all = [int(t.item()) for t in n_all]
        steps = max(n_all)
        if steps == 0:
            return
        cats = torch.full((steps,), -1, dtype=torch.long, device=dev)
        if self._staged:
            cats[:len(self._staged)] = torch.tensor([s[0] for s in self._staged], dtype=torch.long, device=dev)
        cats_all = [torch.zeros_like(cats) for _ in range(world)]
        dist.all_gather(cats_all, cats)
        cats_all = torch.stack(cats_all, dim=1).tolist()  # [step][rank], reference arrival order (:478-485)
        counts = self._counts()
        # local contributions go into zero-initialised deltas so that a repeated sync never re-adds old slots
        d_feats = torch.zeros_like(self.feats_sum)
        d_msum = torch.zeros_like(self.mask_sum)
        d_masks = torch.zeros_like(self.masks)
        for step in range(steps):
            for r in range(world):
                c = cats_all[step][r]
                if c < 0:
                    continue
                slot = counts[c]
                if slot >= self.length:
                    raise IndexError(f"memory bank slot overflow for class {c}")
                if r == rank:
                    _, pooled, wsum, m = self._staged[step]
                    d_feats[c, slot] += pooled
                    d_msum[c, slot] += wsum[0]
                    d_masks[c, slot] += m
                counts[c] += 1
        self._staged = []
        # every slot has exactly one writer: SUM == gather, order-independent and exact for any world size
        flat = torch.cat([d_feats.reshape(-1), d_msum.reshape(-1), d_masks.reshape(-1)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        a = d_feats.numel()
        b = a + d_msum.numel()
        self.feats_sum += flat[:a].view_as(self.feats_sum)
        self.mask_sum += flat[a:b].view_as(self.mask_sum)
        self.masks += flat[b:].view_as(self.masks)
        self.fill_counts.copy_(torch.tensor(counts, dtype=torch.long, device=dev))

    def postprocess(self) -> None:
        """`MemoryBank.postprocess` (`matching_baseline_utils.py:574-656`), the part the test path reads."""
        self.sync_fill()
        ins_avg, avg = ops.fill_finalize(self.feats_sum.contiguous(), self.mask_sum.contiguous())
        self.feats_avg *= 0.0
        self.feats_avg += avg
        self.feats_ins_avg += ins_avg
        self.postprocessed[0] = True
