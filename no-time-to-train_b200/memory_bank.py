"""Memory bank of the matching model, B200-native.

Mirror of `MemoryBank` (`no_time_to_train/models/matching_baseline_utils.py:538-656`): same constructor, same
`postprocess()` entry, same state-dict names for everything the test path reads (`fill_counts`, `masks`,
`feats_avg`, `feats_ins_avg`, `postprocessed`).

What differs, on purpose:
  * the raw `feats [n_cls, L, E, C]` buffer (4.5 GB at 80x10 ViT-L, 67 GB at LVIS scale) is never stored.  Each
    reference shot is reduced on arrival (`nttt_fill_pool_accumulate` / `nttt_fill_pool_batch`) to
    `feats_sum[c,l,:] = sum_e m[e] f[e,:]` and `mask_sum[c,l] = sum_e m[e]`; `postprocess` is then a few KB of
    arithmetic (`nttt_fill_finalize`).
  * multi-GPU fill: the reference all-gathers the raw 5.6 MB feature map of every shot on every step
    (`Sam2MatchingBaseline_noAMG.py:471-474`).  Here every rank pools its own shots straight into a preallocated
    staging table (no allocation, no collective per shot).  `sync_fill()` then (1) all-gathers the tiny per-rank class
    log so that every rank can replay the reference's arrival-order slot assignment (`:478-485`), (2) scatters its
    staged rows into a zero delta buffer with ONE kernel (`nttt_fill_scatter`) and (3) issues ONE
    `all_reduce(SUM)` over `[feats_sum | mask_sum]` (9.8 MB at 80x30x1024; NCCL over NVLink).  Every slot has a
    single writer, so the reduction only adds zeros: the result is bit-identical for any world size.  The low-res
    `masks` buffer (never read after the fill, kept for the state dict) travels in a second, separate all-reduce.
  * `feats_covariances`, `feats_centers`, `ins_sim_avg`, `pca_*` (never read by fill/test, only by the
    out-of-scope `vis_memory`) are not computed.

Checkpoints: POST-PROCESSED checkpoints interchange with the reference in both directions (`strict=False`,
`pl_wrapper/sam2matcher_pl.py:140-142`).  A reference FILL-stage checkpoint (raw `feats`, no `feats_sum`) is reduced
on load (`_load_from_state_dict`).  A fill-stage checkpoint written here has no raw `feats` and cannot be
post-processed by the reference.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops


def _distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class MemoryBank(nn.Module):
    # Lightning's `trainer.save_checkpoint` calls `state_dict()` on EVERY rank (`run_lightning.py:107-119`), so the
    # pending distributed fill may be resolved there.  A rank-local `state_dict()` with shots still staged would
    # deadlock in the collective: set this to False to get a RuntimeError instead and call `sync_fill()` yourself.
    sync_in_state_dict = True
    # None: follow torch.distributed (a fill under an initialised process group of more than one rank is sharded).
    # False: this bank is filled by THIS process alone even under a process group (bench.py's bit-identity check).
    distributed = None

    def _dist(self) -> bool:
        return self.distributed is not False and _distributed()

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
        # staging table of the pending (not yet slotted) shots of THIS rank: rows of pooled sums / mask sums / low-res
        # masks, allocated once and reused; `_stage_cls` is the host-side class log in arrival order
        self._stage_sum = self._stage_wsum = self._stage_mask = None
        self._stage_cls = []
        self.last_sync = {}  # timings / sizes of the most recent sync_fill (read by bench.py)
        self.register_state_dict_pre_hook(MemoryBank._state_dict_hook)

    # ------------------------------------------------------------------------------------------------ state dict
    @staticmethod
    def _state_dict_hook(module, prefix, keep_vars):
        if not module._stage_cls and not module._dist():
            return
        if module.sync_in_state_dict:
            module.sync_fill()
        elif module._stage_cls:
            raise RuntimeError(f"{len(module._stage_cls)} reference shots are staged but not yet slotted: call "
                               "MemoryBank.sync_fill() on every rank before state_dict()")

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        """A reference fill-stage checkpoint carries raw `feats [n_cls, L, E, C]` + `masks` and no `feats_sum`
        (`matching_baseline_utils.py:561-571`): reduce it here exactly as `postprocess` would read it
        (`:581-599`: sum_e feats * masks), class by class so that no second multi-GB tensor appears."""
        raw_key, sum_key = prefix + "feats", prefix + "feats_sum"
        if raw_key in state_dict and sum_key not in state_dict:
            raw, masks = state_dict[raw_key], state_dict.get(prefix + "masks")
            if masks is None:
                error_msgs.append(f"{raw_key} is present but {prefix}masks is not: cannot reduce the raw memory bank")
            elif tuple(raw.shape[:3]) != tuple(masks.shape) or tuple(raw.shape[:2]) != (self.n_classes, self.length) \
                    or raw.shape[-1] != self.feat_shape[1]:
                error_msgs.append(f"{raw_key} has shape {tuple(raw.shape)}, expected "
                                  f"({self.n_classes}, {self.length}, {self.feat_shape[0]}, {self.feat_shape[1]})")
            else:
                sums = torch.empty(raw.shape[0], raw.shape[1], raw.shape[3], dtype=torch.float32, device=raw.device)
                for ci in range(raw.shape[0]):
                    sums[ci] = (raw[ci].float() * masks[ci].float().unsqueeze(-1)).sum(dim=1)
                state_dict = dict(state_dict)
                state_dict[sum_key] = sums
                state_dict[prefix + "mask_sum"] = masks.float().sum(dim=2)
                del state_dict[raw_key]  # consumed (not an unexpected key under strict=True)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        self._host_counts = None

    # ------------------------------------------------------------------------------------------------ fill
    def _counts(self):
        if self._host_counts is None:
            self._host_counts = self.fill_counts.tolist()
        return self._host_counts

    def _stage_rows(self, n_new: int, c: int, e: int, dev):
        """Views of the next `n_new` staging rows (grows the table geometrically; the first allocation already holds
        this rank's share of the whole bank, so a normal fill never reallocates)."""
        used = len(self._stage_cls)
        cap = 0 if self._stage_sum is None else self._stage_sum.shape[0]
        if used + n_new > cap:
            world = dist.get_world_size() if self._dist() else 1
            new_cap = max(used + n_new, 2 * cap, -(-self.n_classes * self.length // world))
            grown = (torch.empty((new_cap, c), dtype=torch.float32, device=dev),
                     torch.empty((new_cap,), dtype=torch.float32, device=dev),
                     torch.empty((new_cap, e), dtype=torch.float32, device=dev))
            if used:
                for new, old in zip(grown, (self._stage_sum, self._stage_wsum, self._stage_mask)):
                    new[:used] = old[:used]
            self._stage_sum, self._stage_wsum, self._stage_mask = grown
        sl = slice(used, used + n_new)
        return self._stage_sum[sl], self._stage_wsum[sl], self._stage_mask[sl]

    def fill(self, cat_ind: int, feat: torch.Tensor, soft_mask: torch.Tensor, enc_hw) -> None:
        """One reference shot (`forward_fill_memory`, `Sam2MatchingBaseline_noAMG.py:465-485`).

        feat [E, C] f32 encoder output, soft_mask [S, S] f32 in [0,1] (nearest-resized to enc_hw inside the
        kernel, like `F.interpolate(mode="nearest")` at :465-469)."""
        feat = feat.reshape(1, -1, feat.shape[-1])
        soft_mask = soft_mask.reshape(1, soft_mask.shape[-2], soft_mask.shape[-1])
        self.fill_batch([int(cat_ind)], feat, soft_mask, enc_hw)

    def fill_batch(self, cat_inds, feats: torch.Tensor, soft_masks: torch.Tensor, enc_hw) -> None:
        """B reference shots in ONE kernel launch: feats [B, E, C] f32, soft_masks [B, S, S] f32, cat_inds B ints (in
        arrival order).  Rows are pooled straight into the staging table; single-process fills are slotted at once,
        distributed fills at `sync_fill()`."""
        cat_inds = [int(c) for c in cat_inds]
        b = len(cat_inds)
        if b == 0:
            return
        if self._dist():
            assert (self.n_classes * self.length) % dist.get_world_size() == 0  # :479-480
        for c in cat_inds:
            if not 0 <= c < self.n_classes:
                raise IndexError(f"class index {c} out of range for a bank of {self.n_classes} classes")
        if len(self._stage_cls) + b > self.n_classes * self.length:
            raise IndexError("memory bank slot overflow: more reference shots than slots")  # reference: IndexError
        dev = self.feats_sum.device
        feats = feats.to(device=dev, dtype=torch.float32).contiguous()
        soft_masks = soft_masks.to(device=dev, dtype=torch.float32).contiguous()
        assert feats.shape[0] == b and soft_masks.shape[0] == b and feats.shape[-1] == self.feat_shape[1]
        e = int(enc_hw[0]) * int(enc_hw[1])
        assert e == self.feat_shape[0] and feats.shape[1] == e
        slots = None if self._dist() else self._assign_slots(cat_inds)  # raises before anything is written
        if slots is not None and b == 1:
            # the model's bs=1 path in a single process: one launch, straight into the slot (no staging, no copies)
            ci, pos = divmod(slots[0], self.length)
            ops.fill_pool_accumulate(feats[0], soft_masks[0], enc_hw, self.feats_sum[ci, pos],
                                     self.mask_sum[ci, pos:pos + 1], self.masks[ci, pos])
            self.fill_counts[ci] += 1
            return
        sums, wsums, masks = self._stage_rows(b, feats.shape[-1], e, dev)
        ops.fill_pool_batch(feats, soft_masks, enc_hw, sums, wsums, masks)
        if slots is None:
            self._stage_cls.extend(cat_inds)
            return
        # single process: the rows go straight from the staging table into their bank slots
        ops.fill_scatter(sums, wsums, masks, torch.tensor(slots, dtype=torch.int32, device=dev), self.feats_sum,
                         self.mask_sum, self.masks)
        self.fill_counts.copy_(torch.tensor(self._counts(), dtype=torch.long))

    def _assign_slots(self, arrival):
        """Replay the reference's slot loop (`:478-485`) over class indices in arrival order; returns the flat slot
        (class * L + position) of each and advances the host counts."""
        counts = list(self._counts())
        slots = []
        for c in arrival:
            if c < 0:
                slots.append(-1)
                continue
            pos = counts[c]
            if pos >= self.length:
                raise IndexError(f"memory bank slot overflow for class {c}")  # the reference raises IndexError too
            slots.append(c * self.length + pos)
            counts[c] += 1
        self._host_counts = counts  # committed only when every shot found a slot
        return slots

    def sync_fill(self) -> None:
        """Resolve the staged distributed fill.  Collective: every rank must call it (no-op without a process group)."""
        if not self._dist():
            assert not self._stage_cls  # single-process fills are slotted on arrival
            return
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = self.feats_sum.device
        n_local = len(self._stage_cls)
        # (1) class log of every rank, fixed size (a rank can never hold more shots than the bank has slots)
        cap = self.n_classes * self.length
        log = torch.full((cap + 1,), -1, dtype=torch.int32)
        log[0] = n_local
        if n_local:
            log[1:1 + n_local] = torch.tensor(self._stage_cls, dtype=torch.int32)
        log = log.to(dev)
        logs = torch.empty((world * (cap + 1),), dtype=torch.int32, device=dev)  # flat: gloo accepts nothing else
        dist.all_gather_into_tensor(logs, log)
        logs = logs.cpu().view(world, cap + 1)
        n_all = logs[:, 0].tolist()
        steps = max(n_all)
        if steps == 0:
            return
        # reference arrival order: step-major, rank-minor (the all_gather + loop of :471-485)
        table = logs[:, 1:1 + steps].t().contiguous().reshape(-1).tolist()
        slots = self._assign_slots(table)
        mine = slots[rank::world][:n_local]
        c, e = self.feat_shape[1], self.feat_shape[0]
        n_slots = self.n_classes * self.length
        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        # (2) local rows -> zero delta [feats_sum | mask_sum], one kernel
        delta = torch.zeros((n_slots * (c + 1),), dtype=torch.float32, device=dev)
        d_feats, d_msum = delta[:n_slots * c].view(n_slots, c), delta[n_slots * c:]
        d_masks = torch.zeros((n_slots, e), dtype=torch.float32, device=dev)
        if n_local:
            slot = torch.tensor(mine, dtype=torch.int32, device=dev)
            ops.fill_scatter(self._stage_sum[:n_local], self._stage_wsum[:n_local], self._stage_mask[:n_local], slot,
                             d_feats, d_msum, d_masks)
        # (3) ONE all-reduce of the sums the test path needs; every slot has exactly one writer: SUM == gather,
        #     order-independent and exact for any world size
        timed = dev.type == "cuda"
        if timed:
            t0.record()
        dist.all_reduce(delta, op=dist.ReduceOp.SUM)
        if timed:
            t1.record()
        dist.all_reduce(d_masks, op=dist.ReduceOp.SUM)  # informational buffer, kept for the state dict
        if timed:
            t2.record()
        self.feats_sum += d_feats.view_as(self.feats_sum)
        self.mask_sum += d_msum.view_as(self.mask_sum)
        self.masks += d_masks.view_as(self.masks)
        self.fill_counts.copy_(torch.tensor(self._counts(), dtype=torch.long))
        self._stage_cls = []
        self.last_sync = dict(world=world, shots=sum(n_all), allreduce_bytes=delta.numel() * 4,
                              masks_allreduce_bytes=d_masks.numel() * 4, class_log_bytes=(cap + 1) * 4,
                              events=(t0, t1, t2) if timed else None)

    # ------------------------------------------------------------------------------------------------ postprocess
    def postprocess(self) -> None:
        """`MemoryBank.postprocess` (`matching_baseline_utils.py:574-656`), the part the test path reads."""
        self.sync_fill()
        if bool((self.masks != 0).any()) and not bool((self.mask_sum != 0).any()):
            raise RuntimeError(
                "memory bank is inconsistent: `masks` holds reference masks but every pooled mask sum is zero — the "
                "state dict that was loaded carried neither `feats_sum` / `mask_sum` nor raw `feats` to reduce them from")
        ins_avg, avg = ops.fill_finalize(self.feats_sum.contiguous(), self.mask_sum.contiguous())
        self.feats_avg *= 0.0
        self.feats_avg += avg
        self.feats_ins_avg += ins_avg
        self.postprocessed[0] = True
