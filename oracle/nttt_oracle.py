"""TEST INFRASTRUCTURE — numpy driver over the plain-C oracle (`oracle/nttt_oracle.c`).

Not the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import
this.  It stitches the C restatements into the reference's matching stage
(`no_time_to_train/models/Sam2MatchingBaseline_noAMG.py:582-683`) with numpy float32 glue, independently
of torch's kernels, so the CUDA path has a checker that shares no code with either torch or the product.
File:line citations are relative to `/root/reference/`.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "nttt_oracle.c")
_LIB = os.path.join(_HERE, "libnttt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (idempotent)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", _SRC, "-o", _LIB, "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_box_nms.restype = ctypes.c_int
        _lib.orc_aa_max_taps.restype = ctypes.c_int
        for fn in (_lib.orc_rle_counts, _lib.orc_rle_to_string, _lib.orc_rle_from_string):
            fn.restype = ctypes.c_long
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def threshold_stats(logits, thr=0.0, off=1.0):
    logits = _f32(logits)
    n, h, w = logits.shape
    mask = np.empty((n, h, w), np.uint8)
    area = np.empty(n, np.int32)
    box = np.empty((n, 4), np.int64)
    hi = np.empty(n, np.int32)
    lo = np.empty(n, np.int32)
    lib().orc_threshold_stats(_p(logits), n, h, w, ctypes.c_float(thr), ctypes.c_float(off), _p(mask),
                              _p(area), _p(box), _p(hi), _p(lo))
    return mask, area, box, hi, lo


def mask_boxes(masks_u8):
    masks_u8 = np.ascontiguousarray(masks_u8, dtype=np.uint8)
    n, h, w = masks_u8.shape
    box = np.empty((n, 4), np.int64)
    area = np.empty(n, np.int32)
    lib().orc_mask_boxes(_p(masks_u8), n, h, w, _p(box), _p(area))
    return box, area


def aa_weights(in_size, out_size):
    taps = lib().orc_aa_max_taps(in_size, out_size)
    xmin = np.empty(out_size, np.int32)
    xsize = np.empty(out_size, np.int32)
    w = np.empty((out_size, taps), np.float32)
    lib().orc_aa_weights(in_size, out_size, taps, _p(xmin), _p(xsize), _p(w))
    return xmin, xsize, w


def aa_resize(src, out_hw):
    src = _f32(src)
    n, ih, iw = src.shape
    dst = np.empty((n, out_hw[0], out_hw[1]), np.float32)
    lib().orc_aa_resize(_p(src), n, ih, iw, out_hw[0], out_hw[1], _p(dst))
    return dst


def aa_resize_threshold(src, out_hw):
    src = _f32(src)
    n, ih, iw = src.shape
    dst = np.empty((n, out_hw[0], out_hw[1]), np.uint8)
    lib().orc_aa_resize_threshold(_p(src), n, ih, iw, out_hw[0], out_hw[1], _p(dst))
    return dst


def pool(masks_u8, feat_cp):
    """masks [N,P] uint8, feat [C,P] -> [N,C] sums (double accumulate)."""
    masks_u8 = np.ascontiguousarray(masks_u8, dtype=np.uint8)
    feat_cp = _f32(feat_cp)
    n, p = masks_u8.shape
    c = feat_cp.shape[0]
    out = np.empty((n, c), np.float32)
    lib().orc_pool(_p(masks_u8), _p(feat_cp), n, p, c, _p(out))
    return out


def box_nms(boxes_f32, scores, labels, thr):
    boxes_f32 = _f32(boxes_f32)
    scores = _f32(scores)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    n = boxes_f32.shape[0]
    keep = np.empty(max(n, 1), np.int64)
    cnt = lib().orc_box_nms(_p(boxes_f32), _p(scores), _p(labels), n, ctypes.c_float(thr), _p(keep))
    return keep[:cnt].copy()


def semantic_ios(masks_u8, labels, obj_sim, want_inter=False):
    masks_u8 = np.ascontiguousarray(masks_u8, dtype=np.uint8)
    k = masks_u8.shape[0]
    flat = masks_u8.reshape(k, -1)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    obj_sim = _f32(obj_sim)
    ios = np.empty(k, np.float32)
    inter = np.empty((k, k), np.int32) if want_inter else None
    lib().orc_semantic_ios(_p(flat), k, ctypes.c_size_t(flat.shape[1]), _p(labels), _p(obj_sim), _p(ios),
                           _p(inter) if want_inter else None)
    return (ios, inter) if want_inter else ios


def select_candidates(multi, ious, iou_thr, first=1):
    """numpy restatement of the candidate selection (Sam2MatchingBaseline_noAMG.py:295-299, :428-431): per prompt the
    first maximal IoU among planes [first, m) (NaN counts as maximal, as torch.argmax), then `score > iou_thr`.
    multi [n, m, h, w], ious [n, m] -> (lr_masks [n', h, w], scores [n'], kept prompt indices [n'])."""
    multi, ious = np.asarray(multi), np.asarray(ious, dtype=np.float32)
    n, m = ious.shape
    best = np.empty(n, dtype=np.int64)
    for i in range(n):
        b, bv = first, ious[i, first]
        for j in range(first + 1, m):
            v = ious[i, j]
            if v > bv or (np.isnan(v) and not np.isnan(bv)):
                b, bv = j, v
        best[i] = b
    scores = ious[np.arange(n), best]
    keep = np.nonzero(scores > np.float32(iou_thr))[0]
    return multi[keep, best[keep]], scores[keep], keep


def rle_counts(mask):
    """pycocotools rleEncode of one [h, w] mask (column-major runs, zeros first) -> uint32 counts."""
    mask = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
    h, w = mask.shape
    cap = 1024
    while True:
        cnts = np.empty(cap, np.uint32)
        m = lib().orc_rle_counts(_p(mask), h, w, _p(cnts), ctypes.c_long(cap))
        if m <= cap:
            return cnts[:m].copy()
        cap = int(m)


def rle_to_string(cnts) -> bytes:
    """pycocotools rleToString."""
    cnts = np.ascontiguousarray(cnts, dtype=np.uint32)
    buf = ctypes.create_string_buffer(7 * max(len(cnts), 1))
    n = lib().orc_rle_to_string(_p(cnts), ctypes.c_long(len(cnts)), buf)
    return buf.raw[:n]


def rle_from_string(s: bytes):
    """pycocotools rleFrString."""
    cnts = np.empty(max(len(s), 1), np.uint32)
    m = lib().orc_rle_from_string(ctypes.c_char_p(s), ctypes.c_long(len(s)), _p(cnts), ctypes.c_long(len(cnts)))
    return cnts[:m].copy()


def rle_decode(cnts, hw):
    """runs -> [h, w] bool mask (column-major), as `rle_to_mask` (sam2/utils/amg.py:143-154)."""
    h, w = hw
    flat = np.zeros(h * w, dtype=bool)
    pos, val = 0, False
    for c in np.asarray(cnts, dtype=np.int64):
        flat[pos:pos + c] = val
        pos += int(c)
        val = not val
    assert pos == h * w
    return flat.reshape(w, h).T


def encode_mask(mask) -> dict:
    """`mask_utils.encode(np.asfortranarray(mask))` with `counts` decoded to str (coco_ref_dataset.py:601-604)."""
    mask = np.asarray(mask)
    return dict(size=[int(mask.shape[0]), int(mask.shape[1])], counts=rle_to_string(rle_counts(mask)).decode("ascii"))


def l2_normalize(x, eps=1e-12):
    """F.normalize(p=2, dim=-1): x / max(||x||, eps)."""
    x = _f32(x)
    nrm = np.sqrt((x.astype(np.float64) ** 2).sum(-1, keepdims=True)).astype(np.float32)
    return x / np.maximum(nrm, np.float32(eps))


def prototypes(feats_ins_avg):
    """normalize(mean over all L slots) (`matching_baseline_utils.py:893-894`)."""
    return l2_normalize(_f32(feats_ins_avg).mean(axis=1, dtype=np.float32))


def match_image(lr_masks, pred_ious, tar_feat, feats_ins_avg, ori_hw, nms_thr=0.5, num_out_instance=100,
                enc_hw=(37, 37), expand_ratio=8):
    """Matching stage for cls_num_per_mask == 1 (every shipped config), numpy + C.

    Follows `Sam2MatchingBaseline_noAMG.py:582-683`; returns the output dict plus intermediates."""
    lr_masks = _f32(lr_masks)
    n, lh, lw = lr_masks.shape
    tar_feat = _f32(tar_feat)
    c = tar_feat.shape[1]
    n_cls = feats_ins_avg.shape[0]
    # :548-549 threshold; amg.py:305-348 boxes; amg.py:158-178 stability counts
    mask_lr, area_lr, box_lr, stab_hi, stab_lo = threshold_stats(lr_masks)
    # :551-558 feature map [C,37,37] -> [C,256,256]
    feat_up = aa_resize(tar_feat.T.reshape(c, enc_hw[0], enc_hw[1]), (lh, lw)).reshape(c, lh * lw)
    # matching_baseline_utils.py:884-891
    sums = pool(mask_lr.reshape(n, -1), feat_up)
    denom = np.where(area_lr == 0, 1, area_lr).astype(np.float32)[:, None]
    obj_feats = l2_normalize(sums / denom)
    sim = obj_feats @ prototypes(feats_ins_avg).T
    # :602-612 with k == 1
    labels = sim.argmax(axis=1).astype(np.int64)
    scores_all = sim[np.arange(n), labels]
    if n_cls == 1:
        scores_all = scores_all * (scores_all > scores_all * np.float32(0.6))
    # :621-629
    out_num = int(min(num_out_instance * expand_ratio, n))
    keep = box_nms(box_lr.astype(np.float32), pred_ious, labels, nms_thr)[:out_num]
    # :631-641
    pos = scores_all[keep] > 0
    sel = keep[pos]
    res = dict(sim=sim, obj_feats=obj_feats, labels_all=labels, scores_all=scores_all, lr_boxes=box_lr,
               lr_area=area_lr, stab_hi=stab_hi, stab_lo=stab_lo, keep=keep, sel_index=sel)
    oh, ow = ori_hw
    if sel.shape[0] == 0:
        res.update(binary_masks=np.zeros((0, oh, ow), np.uint8), bboxes=np.zeros((0, 4), np.float32),
                   scores=np.zeros(0, np.float32), labels=np.zeros(0, np.int64))
        return res
    # :657-665
    full = aa_resize_threshold(lr_masks[sel], (oh, ow))
    full_boxes, full_area = mask_boxes(full)
    # :668-672
    f_sel = obj_feats[sel]
    obj_sim = np.maximum(f_sel @ f_sel.T, np.float32(0))
    ios, inter = semantic_ios(full, labels[sel], obj_sim, want_inter=True)
    with np.errstate(invalid="ignore"):
        decayed = scores_all[sel] * np.sqrt(np.float32(1) - ios)
    # :674-675 argsort(descending=True): NaN first, then by value; ties -> lower index (stable)
    n_out = min(num_out_instance, decayed.shape[0])
    key = np.where(np.isnan(decayed), np.float32(np.inf), decayed)
    order = np.argsort(-key, kind="stable")[:n_out]
    res.update(binary_masks=full[order], bboxes=full_boxes[order], scores=decayed[order],
               labels=labels[sel][order], ios=ios, inter=inter, decayed=decayed, full_area=full_area,
               full_boxes=full_boxes, order=order, labels_sel=labels[sel])
    return res
