"""TEST INFRASTRUCTURE — torch restatement ("port") of the reference's matching stage.

This file is the oracle, not the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
CPU-baseline / `--impl reference` legs may import it.  It restates, with the same library calls the
reference makes (aten `_upsample_bilinear2d_aa`, SGEMM, torchvision `batched_nms`, `topk`, `argsort`),
the part of `Sam2MatchingBaselineNoAMG.forward_test` that runs after the frozen encoders, plus the
memory-bank fill / post-process.  File:line citations are relative to `/root/reference/`.

Pinning: `tests/golden/make_golden.py` executes the REAL reference functions (imported through
`tests/golden/ref_shim.py`) on seeded inputs and commits their outputs under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this restatement (and the C/numpy oracle in `oracle/nttt_oracle.py`)
against those vectors.  The reference itself ships no tests or fixtures for this path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.nn.functional as F
from torchvision.ops.boxes import batched_nms


@dataclass
class StageConfig:
    """The `sam2_infer_cfgs` constants the stage reads (`Sam2MatchingBaseline_noAMG.py:185-194`)."""
    nms_thr: float = 0.5
    num_out_instance: int = 100
    cls_num_per_mask: int = 1
    enc_hw: tuple = (37, 37)
    expand_ratio: int = 8  # literal at :621


def select_candidates(chunks, ious_chunks, iou_thr: float):
    """Candidate selection in front of the stage: per decoder batch `best = argmax(ious[:, 1:]) + 1`, gather that
    plane and its IoU (`_forward_sam_decoder`, :295-299), concatenate the batches (:423-425), keep
    `scores > iou_thr` (:428-431).  chunks: list of [bs, m, h, w]; ious_chunks: list of [bs, m].
    -> lr_masks [N', h, w], pred_ious [N'], kept prompt indices [N']."""
    masks, scores = [], []
    for multi, ious in zip(chunks, ious_chunks):
        best = torch.argmax(ious[:, 1:], dim=-1) + 1
        rows = torch.arange(multi.shape[0])
        masks.append(multi[rows, best])
        scores.append(ious[rows, best].reshape(-1))
    masks = torch.cat(masks, dim=0)
    scores = torch.cat(scores, dim=0).reshape(-1)
    inds = scores > iou_thr
    return masks[inds], scores[inds], torch.nonzero(inds).reshape(-1)


def threshold_lowres(lr_masks: torch.Tensor) -> torch.Tensor:
    """`_process_sam_masks`, mask half (:547-549): strict `> 0`, flattened to [N, P]."""
    return (lr_masks > 0).reshape(lr_masks.shape[0], -1)


def upsample_features(tar_feat: torch.Tensor, enc_hw, out_hw) -> torch.Tensor:
    """`_process_sam_masks`, feature half (:551-558): [E,C] -> antialiased bilinear -> [P,C] view."""
    eh, ew = enc_hw
    spatial = tar_feat.reshape(1, eh, ew, -1).permute(0, 3, 1, 2)
    up = F.interpolate(spatial, size=tuple(out_hw), mode="bilinear", align_corners=False, antialias=True)
    return up.reshape(-1, out_hw[0] * out_hw[1]).t()


def prototypes(feats_ins_avg: torch.Tensor) -> torch.Tensor:
    """Class prototypes: mean over ALL L slots (unfilled zeros included), then L2-normalise
    (`matching_baseline_utils.py:893-894`)."""
    return F.normalize(feats_ins_avg.mean(dim=1), p=2, dim=-1)


def pool_and_score(feat_pc: torch.Tensor, masks_bool: torch.Tensor, feats_ins_avg: torch.Tensor):
    """`compute_sim_global_avg(..., softmax=False, temp=1.0, ret_feats=True)`
    (`matching_baseline_utils.py:869-904`)."""
    m = masks_bool.to(feat_pc.dtype)
    area = m.sum(dim=-1, keepdim=True)
    area[area == 0] = 1.0
    pooled = F.normalize((m @ feat_pc) / area, p=2, dim=-1)
    sim = pooled @ prototypes(feats_ins_avg).t()
    return sim / 1.0, pooled


def pool_and_score_neg(feat_pc, masks_bool, feats_avg, feats_ins_avg_neg, sigma=0.8):
    """`compute_sim_global_avg_with_neg` (`matching_baseline_utils.py:906-941`) + the obj_feats the caller
    recomputes (`Sam2MatchingBaseline_noAMG.py:597-600`).  No zero guard on the area here: empty masks give NaN."""
    n = masks_bool.shape[0]
    n_cls, c = feats_avg.shape
    m = masks_bool.to(feat_pc.dtype)
    pooled = F.normalize((m @ feat_pc) / m.sum(dim=-1, keepdim=True), p=2, dim=-1)
    pos = F.normalize(feats_avg, p=2, dim=-1)
    neg = F.normalize(feats_ins_avg_neg, p=2, dim=-1).reshape(-1, c)
    sim_pos = (pooled @ pos.t()).clamp(min=0.0)
    sim_neg = (pooled @ neg.t()).clamp(min=0.0).reshape(n, n_cls, -1).max(dim=-1).values
    sim = sim_pos * torch.exp(-1.0 * (sim_neg - sim_pos).clamp(min=0.0) / sigma)
    return sim, pooled


def select_labels(sim: torch.Tensor, k: int):
    """top-k / label section (`Sam2MatchingBaseline_noAMG.py:602-612`)."""
    n_cls = sim.shape[1]
    if k == -1:
        k = n_cls
    top, labels = torch.topk(sim, k=k)
    if k == n_cls:
        top = top * (top > (top[:, 0:1] * 0.6))
    return top.flatten(), labels.flatten(), k


def mask_boxes(masks_bool: torch.Tensor) -> torch.Tensor:
    """`batched_mask_to_box` (`sam2/utils/amg.py:305-348`): XYXY, inclusive max index, empty -> zeros,
    int64.  Restated with any/argmax-free reductions."""
    n, h, w = masks_bool.shape
    if masks_bool.numel() == 0:
        return torch.zeros(n, 4, device=masks_bool.device)
    rows = masks_bool.any(dim=2)
    cols = masks_bool.any(dim=1)
    ar_h = torch.arange(h, device=masks_bool.device)
    ar_w = torch.arange(w, device=masks_bool.device)
    bottom = (rows * ar_h).max(dim=1).values
    top = (rows * ar_h + h * (~rows)).min(dim=1).values
    right = (cols * ar_w).max(dim=1).values
    left = (cols * ar_w + w * (~cols)).min(dim=1).values
    empty = (right < left) | (bottom < top)
    out = torch.stack([left, top, right, bottom], dim=-1)
    return out * (~empty).unsqueeze(-1)


def stability_score(logits: torch.Tensor, thr: float, off: float) -> torch.Tensor:
    """`calculate_stability_score` (`sam2/utils/amg.py:158-178`): int counts, true-divide (0/0 -> NaN)."""
    hi = (logits > (thr + off)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    lo = (logits > (thr - off)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    return hi / lo


def upsample_threshold(lr_sel: torch.Tensor, ori_hw) -> torch.Tensor:
    """final resize + threshold (`Sam2MatchingBaseline_noAMG.py:657-663`)."""
    up = F.interpolate(lr_sel.unsqueeze(1), size=tuple(ori_hw), mode="bilinear", align_corners=False,
                       antialias=True)
    return up.squeeze(1) > 0


def semantic_ios(masks_bool: torch.Tensor, labels: torch.Tensor, obj_sim: torch.Tensor, n_cls: int):
    """`compute_semantic_ios(use_semantic=True, rank_score=True)` (`matching_baseline_utils.py:831-867`).

    Same arithmetic order as the reference: (inter * sim) / area_i * sim, diagonal zeroed, row max."""
    k = masks_bool.shape[0]
    flat = masks_bool.reshape(k, -1).float()
    ios = torch.zeros(k, dtype=torch.float32, device=masks_bool.device)
    for c in labels.unique().tolist() if k else []:
        if c < 0 or c >= n_cls:
            continue
        sel = labels == c
        grp = flat[sel]
        s = obj_sim[sel][:, sel]
        area = grp.sum(dim=-1)
        inter = grp @ grp.t()
        inter.fill_diagonal_(0.0)
        val = ((inter * s) / area[:, None]) * s
        ios[sel] += val.max(dim=-1).values
    return ios


def match_image(lr_masks, pred_ious, tar_feat, feats_ins_avg, cfg: StageConfig, ori_hw, timings=None,
                override=None, negative=None):
    """The matching stage of `forward_test(with_negative=False)`
    (`Sam2MatchingBaseline_noAMG.py:582-683`), from the `_forward_sam` seam to the output dict.

    `timings`, if a dict, receives per-section wall-clock seconds (CPU baseline reporting).
    `override`, if a dict with `sim` and `obj_feats`, replaces the pooled features / similarities so that a
    test can check everything downstream of the float contractions exactly (label near-ties otherwise
    cascade through NMS)."""
    import time

    def lap(name, t0):
        if timings is not None:
            timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    t = time.perf_counter()
    device = lr_masks.device
    n_cls = feats_ins_avg.shape[0] if negative is None else negative["feats_avg"].shape[0]
    masks_bool = threshold_lowres(lr_masks)
    feat_pc = upsample_features(tar_feat, cfg.enc_hw, lr_masks.shape[-2:])
    t = lap("process_sam_masks", t)
    if negative is not None:  # dict(feats_avg=..., feats_ins_avg_neg=..., sigma=0.8): with_negative=True path
        sim, obj_feats = pool_and_score_neg(feat_pc, masks_bool, negative["feats_avg"], negative["feats_ins_avg_neg"],
                                            negative.get("sigma", 0.8))
    else:
        sim, obj_feats = pool_and_score(feat_pc, masks_bool, feats_ins_avg)
    if override is not None:
        sim, obj_feats = override["sim"], override["obj_feats"]
    t = lap("pool_and_score", t)
    scores_all, labels, k = select_labels(sim, cfg.cls_num_per_mask)
    t = lap("topk", t)
    lr_boxes = mask_boxes(lr_masks > 0)
    boxes_exp = lr_boxes.unsqueeze(1).expand(-1, k, -1).reshape(lr_masks.shape[0] * k, 4)
    t = lap("lowres_boxes", t)
    out_num = int(min(cfg.num_out_instance * cfg.expand_ratio, labels.shape[0]))
    keep = batched_nms(boxes_exp.float(), pred_ious.flatten(), labels, iou_threshold=cfg.nms_thr)[:out_num]
    t = lap("batched_nms", t)

    scores = scores_all[keep]
    lr_sel = lr_masks[keep // k]
    feats_sel = obj_feats[keep // k]
    labels_sel = labels[keep]
    pos = scores > 0.0
    scores, lr_sel, feats_sel, labels_sel = scores[pos], lr_sel[pos], feats_sel[pos], labels_sel[pos]
    sel_index = (keep // k)[pos]
    t = lap("gather", t)

    oh, ow = ori_hw
    if lr_sel.shape[0] == 0:
        return dict(
            binary_masks=torch.zeros((0, oh, ow), device=device, dtype=torch.bool),
            bboxes=torch.zeros((0, 4), device=device, dtype=torch.float32),
            scores=torch.zeros((0,), device=device, dtype=torch.float32),
            labels=torch.zeros((0,), device=device, dtype=torch.long),
            aux=dict(sim=sim, obj_feats=obj_feats, labels_all=labels, scores_all=scores_all,
                     lr_boxes=lr_boxes, keep=keep, sel_index=sel_index),
        )

    full = upsample_threshold(lr_sel, (oh, ow))
    t = lap("upsample_threshold", t)
    boxes = mask_boxes(full)
    t = lap("fullres_boxes", t)
    obj_sim = (feats_sel @ feats_sel.t()).clamp(min=0.0)
    ios = semantic_ios(full, labels_sel, obj_sim, n_cls)
    t = lap("semantic_ios", t)
    decayed = scores * torch.pow(1 - ios, 0.5)
    n_out = min(cfg.num_out_instance, decayed.shape[0])
    order = torch.argsort(decayed, descending=True)[:n_out]
    out = dict(
        binary_masks=full[order],
        bboxes=boxes[order],
        scores=decayed[order],
        labels=labels_sel[order],
        aux=dict(sim=sim, obj_feats=obj_feats, labels_all=labels, scores_all=scores_all,
                 lr_boxes=lr_boxes, keep=keep, sel_index=sel_index, ios=ios, decayed=decayed,
                 full_area=full.reshape(full.shape[0], -1).sum(-1), full_boxes=boxes, order=order),
    )
    lap("decay_topk", t)
    return out


# ----------------------------------------------------------------------------------------------------
# memory bank
# ----------------------------------------------------------------------------------------------------

@dataclass
class RawBank:
    """The raw buffers of the reference `MemoryBank` (`matching_baseline_utils.py:561-571`)."""
    n_cls: int
    length: int
    e: int
    c: int
    feats: torch.Tensor = field(init=False)
    masks: torch.Tensor = field(init=False)
    fill_counts: torch.Tensor = field(init=False)

    def __post_init__(self):
        self.feats = torch.zeros(self.n_cls, self.length, self.e, self.c)
        self.masks = torch.zeros(self.n_cls, self.length, self.e)
        self.fill_counts = torch.zeros(self.n_cls, dtype=torch.long)


def bank_fill(bank: RawBank, cat_inds, feats, masks) -> None:
    """The slot loop of `forward_fill_memory` (`Sam2MatchingBaseline_noAMG.py:478-485`) over the
    gathered samples, in arrival order."""
    for i in range(len(cat_inds)):
        c = int(cat_inds[i])
        slot = int(bank.fill_counts[c])
        bank.feats[c, slot] += feats[i]
        bank.masks[c, slot] += masks[i]
        bank.fill_counts[c] += 1


def bank_postprocess(bank: RawBank):
    """`MemoryBank.postprocess` (`matching_baseline_utils.py:574-599`), the two outputs read at test
    time: feats_avg [n_cls,C] and feats_ins_avg [n_cls,L,C]."""
    w_all = bank.masks.sum(dim=(1, 2)).unsqueeze(1)
    w_all[w_all == 0] = 1.0
    weighted = bank.feats * bank.masks.unsqueeze(-1)
    feats_avg = weighted.sum(dim=(1, 2)) / w_all
    w_ins = bank.masks.sum(dim=2).unsqueeze(2)
    w_ins[w_ins == 0] = 1.0
    feats_ins_avg = weighted.sum(dim=2) / w_ins
    return feats_avg, feats_ins_avg
