"""Per-operator Python entry points over the C-ABI (one function per `nttt_*` entry).

torch is used for device memory and streams only: every function allocates its outputs with torch, passes
raw device pointers and the current CUDA stream to libnttt_b200.so, and returns torch tensors.  Nothing here
synchronises the device.  All inputs must be contiguous CUDA tensors of the documented dtype.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_ctx_by_device = {}


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the matching stage has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def context(device) -> int:
    """The per-device `nttt_ctx*` (created on first use)."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ctx_by_device:
        lib = _lib.load()
        out = ctypes.c_void_p()
        with torch.cuda.device(idx):
            _lib.check(lib.nttt_ctx_create(ctypes.byref(out), idx), "nttt_ctx_create")
        _ctx_by_device[idx] = out.value
    return _ctx_by_device[idx]


def tune(device, what: int, value: int) -> None:
    """`nttt_ctx_tune` on the device's context (include/nttt_b200.h NTTT_TUNE_*)."""
    _lib.check(_lib.load().nttt_ctx_tune(context(device), int(what), int(value)), "nttt_ctx_tune")


def threshold_pack(logits: torch.Tensor, thr: float = 0.0, off: float = 1.0, out=None, want_stab: bool = True):
    """-> bits [n, h*w/32] int32 (bit pattern of uint32), area [n], box [n,4], stab [n,2], flags [n].
    `out` may carry the five preallocated outputs of an earlier call (no allocation inside a timed loop).
    want_stab=False skips the stability counts (the variant `nttt_match_image` launches)."""
    _need(logits, torch.float32, "logits")
    n, h, w = logits.shape
    dev = logits.device
    if out is not None:
        bits, area, box, stab, flags = out
    else:
        bits = torch.empty((n, h * w // 32), dtype=torch.int32, device=dev)
        area = torch.empty((n,), dtype=torch.int32, device=dev)
        box = torch.empty((n, 4), dtype=torch.int32, device=dev)
        stab = torch.empty((n, 2), dtype=torch.int32, device=dev)
        flags = torch.empty((n,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    _lib.check(lib.nttt_threshold_pack(_ptr(logits), n, h, w, thr, off, _ptr(bits), _ptr(area), _ptr(box),
                                       _ptr(stab) if want_stab else None, _ptr(flags), _stream(dev)),
               "nttt_threshold_pack")
    return bits, area, box, stab, flags


def calculate_stability_score(logits: torch.Tensor, mask_threshold: float, threshold_offset: float) -> torch.Tensor:
    """`calculate_stability_score(masks, mask_threshold, threshold_offset)` (`sam2/utils/amg.py:158-178`) — same name,
    arguments and result: count(logit > thr+off) / count(logit > thr-off) per mask as float32 (0/0 -> NaN), computed
    in the same single pass over the logits that packs the masks (`nttt_threshold_pack_stability`).
    logits [..., h, w] f32 -> [...] f32."""
    _need(logits, torch.float32, "logits")
    lead, (h, w) = logits.shape[:-2], logits.shape[-2:]
    flat = logits.reshape(-1, h, w)
    n = flat.shape[0]
    dev = logits.device
    bits = torch.empty((n, h * w // 32), dtype=torch.int32, device=dev)
    area = torch.empty((n,), dtype=torch.int32, device=dev)
    box = torch.empty((n, 4), dtype=torch.int32, device=dev)
    stab = torch.empty((n, 2), dtype=torch.int32, device=dev)
    flags = torch.empty((n,), dtype=torch.int32, device=dev)
    score = torch.empty((n,), dtype=torch.float32, device=dev)
    lib = _lib.load()
    _lib.check(lib.nttt_threshold_pack_stability(_ptr(flat), n, h, w, float(mask_threshold), float(threshold_offset),
                                                 _ptr(bits), _ptr(area), _ptr(box), _ptr(stab), _ptr(score), _ptr(flags),
                                                 _stream(dev)), "nttt_threshold_pack_stability")
    return score.reshape(lead)


def chunk_table(chunks):
    """HOST array of the device pointers of the decoder's per-batch tensors (each [prompts, m, h, w] f32)."""
    arr = (ctypes.c_void_p * len(chunks))()
    for i, t in enumerate(chunks):
        _need(t, torch.float32, "logits chunk")
        arr[i] = t.data_ptr()
    return arr


def select_multimask(ious: torch.Tensor, chunks, first: int = 1):
    """Best decoder plane per prompt (`argmax(ious[:, first:]) + first`, `Sam2MatchingBaseline_noAMG.py:295-299`).
    `chunks`: the decoder's per-batch outputs (list of [prompts, m, h, w] tensors, equal prompts except the last).
    -> mask_ptr [n] int64 (device addresses of the chosen planes), score [n] f32."""
    _need(ious, torch.float32, "ious")
    n, m = ious.shape
    _, m2, h, w = chunks[0].shape
    assert m2 == m
    mask_ptr = torch.empty((n,), dtype=torch.int64, device=ious.device)
    score = torch.empty((n,), dtype=torch.float32, device=ious.device)
    lib = _lib.load()
    _lib.check(lib.nttt_select_multimask(_ptr(ious), n, m, first, chunk_table(chunks), len(chunks), chunks[0].shape[0],
                                         h, w, _ptr(mask_ptr), _ptr(score), _stream(ious.device)),
               "nttt_select_multimask")
    return mask_ptr, score


def threshold_pack_ptrs(mask_ptr: torch.Tensor, hw, gate=None, gate_min: float = 0.0, thr: float = 0.0,
                        off: float = 1.0):
    """`threshold_pack` of mask i read from address mask_ptr[i]; masks with !(gate[i] > gate_min) are published as
    empty without reading their logits."""
    _need(mask_ptr, torch.int64, "mask_ptr")
    h, w = hw
    n = mask_ptr.shape[0]
    dev = mask_ptr.device
    bits = torch.empty((n, h * w // 32), dtype=torch.int32, device=dev)
    area = torch.empty((n,), dtype=torch.int32, device=dev)
    box = torch.empty((n, 4), dtype=torch.int32, device=dev)
    stab = torch.empty((n, 2), dtype=torch.int32, device=dev)
    flags = torch.empty((n,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    _lib.check(lib.nttt_threshold_pack_ptrs(_ptr(mask_ptr), _ptr(gate), gate_min, n, h, w, thr, off, _ptr(bits),
                                            _ptr(area), _ptr(box), _ptr(stab), _ptr(flags), _stream(dev)),
               "nttt_threshold_pack_ptrs")
    return bits, area, box, stab, flags


def project_masks(bits: torch.Tensor, box: torch.Tensor, hw, enc_hw):
    _need(bits, torch.int32, "bits")
    _need(box, torch.int32, "box")
    n = bits.shape[0]
    proj = torch.empty((n, enc_hw[0] * enc_hw[1]), dtype=torch.float32, device=bits.device)
    lib = _lib.load()
    _lib.check(lib.nttt_project_masks(context(bits.device), _ptr(bits), _ptr(box), n, hw[0], hw[1], enc_hw[0], enc_hw[1],
                                      _ptr(proj), _stream(bits.device)), "nttt_project_masks")
    return proj


def pool_normalize(proj: torch.Tensor, feat: torch.Tensor, area: torch.Tensor):
    _need(proj, torch.float32, "proj")
    _need(feat, torch.float32, "feat")
    _need(area, torch.int32, "area")
    n, e = proj.shape
    c = feat.shape[1]
    lib = _lib.load()
    ws_bytes = lib.nttt_pool_workspace_bytes(n, e, c)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=proj.device)
    out = torch.empty((n, c), dtype=torch.float32, device=proj.device)
    _lib.check(lib.nttt_pool_normalize(context(proj.device), _ptr(proj), _ptr(feat), _ptr(area), n, e, c, _ptr(out),
                                       _ptr(ws), ws_bytes, _stream(proj.device)), "nttt_pool_normalize")
    return out


def proto_prepare(feats_ins_avg: torch.Tensor):
    _need(feats_ins_avg, torch.float32, "feats_ins_avg")
    n_cls, shots, c = feats_ins_avg.shape
    proto = torch.empty((n_cls, c), dtype=torch.float32, device=feats_ins_avg.device)
    lib = _lib.load()
    _lib.check(lib.nttt_proto_prepare(_ptr(feats_ins_avg), n_cls, shots, c, _ptr(proto),
                                      _stream(feats_ins_avg.device)), "nttt_proto_prepare")
    return proto


def similarity_top1(obj_feats: torch.Tensor, proto: torch.Tensor, want_sim: bool = True):
    _need(obj_feats, torch.float32, "obj_feats")
    _need(proto, torch.float32, "proto")
    n, c = obj_feats.shape
    n_cls = proto.shape[0]
    dev = obj_feats.device
    lib = _lib.load()
    sim = torch.empty((n, n_cls), dtype=torch.float32, device=dev) if want_sim else None
    ws_bytes = lib.nttt_similarity_workspace_bytes(n, c, n_cls)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
    top_score = torch.empty((n,), dtype=torch.float32, device=dev)
    top_label = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.check(lib.nttt_similarity_top1(context(dev), _ptr(obj_feats), _ptr(proto), n, c, n_cls, _ptr(sim),
                                        _ptr(top_score), _ptr(top_label), _ptr(ws), ws_bytes, _stream(dev)),
               "nttt_similarity_top1")
    return sim, top_score, top_label


def similarity_neg_top1(obj_feats: torch.Tensor, proto_pos: torch.Tensor, proto_neg: torch.Tensor, l_neg: int,
                        sigma: float = 0.8):
    """Negative-reference scoring: proto_pos [n_cls, c], proto_neg [n_cls*l_neg, c], all unit rows."""
    _need(obj_feats, torch.float32, "obj_feats")
    _need(proto_pos, torch.float32, "proto_pos")
    _need(proto_neg, torch.float32, "proto_neg")
    n, c = obj_feats.shape
    n_cls = proto_pos.shape[0]
    dev = obj_feats.device
    lib = _lib.load()
    sim = torch.empty((n, n_cls), dtype=torch.float32, device=dev)
    ws_bytes = lib.nttt_similarity_neg_workspace_bytes(n, c, n_cls, l_neg)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
    top_score = torch.empty((n,), dtype=torch.float32, device=dev)
    top_label = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.check(lib.nttt_similarity_neg_top1(context(dev), _ptr(obj_feats), _ptr(proto_pos), _ptr(proto_neg), n, c, n_cls,
                                            l_neg, sigma, _ptr(sim), _ptr(top_score), _ptr(top_label), _ptr(ws),
                                            ws_bytes, _stream(dev)), "nttt_similarity_neg_top1")
    return sim, top_score, top_label


def box_nms(box: torch.Tensor, nms_scores: torch.Tensor, labels: torch.Tensor, top_score: torch.Tensor,
            iou_thr: float, max_keep: int):
    """-> keep [max_keep] i32, sel [max_keep] i32, counts [2] i32 = (n_keep, n_sel); all on device."""
    _need(box, torch.int32, "box")
    _need(nms_scores, torch.float32, "nms_scores")
    _need(labels, torch.int32, "labels")
    _need(top_score, torch.float32, "top_score")
    n = box.shape[0]
    dev = box.device
    lib = _lib.load()
    ws_bytes = lib.nttt_nms_workspace_bytes(n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    keep = torch.zeros((max(max_keep, 1),), dtype=torch.int32, device=dev)
    sel = torch.zeros((max(max_keep, 1),), dtype=torch.int32, device=dev)
    counts = torch.zeros((2,), dtype=torch.int32, device=dev)
    _lib.check(lib.nttt_box_nms(_ptr(box), _ptr(nms_scores), _ptr(labels), _ptr(top_score), n, iou_thr, max_keep,
                                _ptr(keep), counts.data_ptr(), _ptr(sel), counts.data_ptr() + 4, _ptr(ws), ws_bytes,
                                _stream(dev)), "nttt_box_nms")
    return keep, sel, counts


def upsample_threshold_pack(logits, bits_lr, box_lr, flags_lr, sel, n_sel, max_sel: int, ori_hw, mask_ptr=None,
                            lr_hw=None):
    """-> bits_full [max_sel, oh, words] i32, rect [max_sel,4], area_full [max_sel], box_full [max_sel,4].
    With `mask_ptr` (from select_multimask) candidate i's logits are read from mask_ptr[i]; pass logits=None and
    lr_hw=(ih, iw)."""
    _need(bits_lr, torch.int32, "bits_lr")
    _need(sel, torch.int32, "sel")
    _need(n_sel, torch.int32, "n_sel")
    if mask_ptr is not None:
        ih, iw = lr_hw
        logits = bits_lr  # device handle only
    else:
        _need(logits, torch.float32, "logits")
        _, ih, iw = logits.shape
    oh, ow = ori_hw
    dev = logits.device
    words = (ow + 31) // 32
    bits_full = torch.empty((max_sel, oh, words), dtype=torch.int32, device=dev)
    rect = torch.zeros((max_sel, 4), dtype=torch.int32, device=dev)
    area_full = torch.zeros((max_sel,), dtype=torch.int32, device=dev)
    box_full = torch.zeros((max_sel, 4), dtype=torch.int32, device=dev)
    lib = _lib.load()
    if mask_ptr is not None:
        _lib.check(lib.nttt_upsample_threshold_pack_ptrs(context(dev), _ptr(mask_ptr), _ptr(bits_lr), _ptr(box_lr),
                                                         _ptr(flags_lr), ih, iw, _ptr(sel), _ptr(n_sel), max_sel, oh, ow,
                                                         _ptr(bits_full), _ptr(rect), _ptr(area_full), _ptr(box_full),
                                                         _stream(dev)), "nttt_upsample_threshold_pack_ptrs")
        return bits_full, rect, area_full, box_full
    _lib.check(lib.nttt_upsample_threshold_pack(context(dev), _ptr(logits), _ptr(bits_lr), _ptr(box_lr),
                                                _ptr(flags_lr), ih, iw, _ptr(sel), _ptr(n_sel), max_sel, oh, ow,
                                                _ptr(bits_full), _ptr(rect), _ptr(area_full), _ptr(box_full),
                                                _stream(dev)), "nttt_upsample_threshold_pack")
    return bits_full, rect, area_full, box_full


def unpack_masks(bits_full, rect, n_sel, ori_hw):
    max_sel = bits_full.shape[0]
    oh, ow = ori_hw
    out = torch.zeros((max_sel, oh, ow), dtype=torch.uint8, device=bits_full.device)
    lib = _lib.load()
    _lib.check(lib.nttt_unpack_masks(_ptr(bits_full), _ptr(rect), _ptr(n_sel), max_sel, oh, ow, _ptr(out),
                                     _stream(bits_full.device)), "nttt_unpack_masks")
    return out.view(torch.bool)


def mask_ios(bits_full, rect, area_full, box_full, sel, n_sel, ori_hw, labels, obj_feats, want_inter=False):
    max_sel = bits_full.shape[0]
    oh, ow = ori_hw
    dev = bits_full.device
    c = obj_feats.shape[1]
    ios = torch.zeros((max_sel,), dtype=torch.float32, device=dev)
    inter = torch.zeros((max_sel, max_sel), dtype=torch.int32, device=dev) if want_inter else None
    lib = _lib.load()
    ws_bytes = lib.nttt_mask_ios_workspace_bytes(max_sel)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    _lib.check(lib.nttt_mask_ios(_ptr(bits_full), _ptr(rect), _ptr(area_full), _ptr(box_full), _ptr(sel), _ptr(n_sel),
                                 max_sel, oh, ow, _ptr(labels), _ptr(obj_feats), c, _ptr(ios), _ptr(inter), _ptr(ws),
                                 ws_bytes, _stream(dev)), "nttt_mask_ios")
    return (ios, inter) if want_inter else ios


def decay_topk(top_score, labels, ios, sel, n_sel, num_out: int, bits_full, rect, box_full, ori_hw):
    max_sel = bits_full.shape[0]
    oh, ow = ori_hw
    dev = bits_full.device
    out_masks = torch.empty((num_out, oh, ow), dtype=torch.uint8, device=dev)
    out_boxes = torch.zeros((num_out, 4), dtype=torch.int64, device=dev)
    out_scores = torch.zeros((num_out,), dtype=torch.float32, device=dev)
    out_labels = torch.zeros((num_out,), dtype=torch.int64, device=dev)
    out_index = torch.zeros((num_out,), dtype=torch.int32, device=dev)
    out_slot = torch.zeros((num_out,), dtype=torch.int32, device=dev)
    n_out = torch.zeros((1,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    _lib.check(lib.nttt_decay_topk(_ptr(top_score), _ptr(labels), _ptr(ios), _ptr(sel), _ptr(n_sel), max_sel, num_out,
                                   _ptr(bits_full), _ptr(rect), _ptr(box_full), oh, ow, _ptr(out_masks),
                                   _ptr(out_boxes), _ptr(out_scores), _ptr(out_labels), _ptr(out_index),
                                   _ptr(out_slot), _ptr(n_out), _stream(dev)), "nttt_decay_topk")
    return out_masks.view(torch.bool), out_boxes, out_scores, out_labels, out_index, out_slot, n_out


def rle_encode(bits_full, rect, slot, count, ori_hw, cap_counts: int = 16384, cap_chars: int = 32768, max_count=None):
    """COCO compressed RLE of packed masks (pycocotools `mask_utils.encode` wire format, column-major).
    -> counts [max_count, cap_counts] int32 (uint32 bit pattern), n_counts [max_count], chars [max_count, cap_chars]
    uint8, n_chars [max_count]."""
    _need(bits_full, torch.int32, "bits_full")
    _need(rect, torch.int32, "rect")
    _need(count, torch.int32, "count")
    oh, ow = ori_hw
    dev = bits_full.device
    if max_count is None:
        max_count = slot.shape[0] if slot is not None else bits_full.shape[0]
    counts = torch.empty((max_count, cap_counts), dtype=torch.int32, device=dev)
    n_counts = torch.empty((max_count,), dtype=torch.int32, device=dev)
    chars = torch.empty((max_count, cap_chars), dtype=torch.uint8, device=dev)
    n_chars = torch.empty((max_count,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    _lib.check(lib.nttt_rle_encode(_ptr(bits_full), _ptr(rect), _ptr(slot), _ptr(count), max_count, oh, ow, cap_counts,
                                   cap_chars, _ptr(counts), _ptr(n_counts), _ptr(chars), _ptr(n_chars), _stream(dev)),
               "nttt_rle_encode")
    return counts, n_counts, chars, n_chars


def rle_compact(chars, n_chars, n_masks: int, out=None):
    """The strings of outputs 0..n_masks-1 packed back to back (`nttt_rle_compact`): -> uint8 [>= sum(len)] on the
    device; mask j starts at sum_{i<j} len_i.  Lengths outside [0, cap_chars] (overflowed masks) count as 0."""
    _need(chars, torch.uint8, "chars")
    _need(n_chars, torch.int32, "n_chars")
    if out is None:
        out = torch.empty((max(n_masks, 1) * chars.shape[1],), dtype=torch.uint8, device=chars.device)
    _need(out, torch.uint8, "out")
    lib = _lib.load()
    _lib.check(lib.nttt_rle_compact(_ptr(chars), _ptr(n_chars), int(n_masks), int(chars.shape[1]), _ptr(out),
                                    out.numel(), _stream(chars.device)), "nttt_rle_compact")
    return out


def fill_pool_accumulate(feat, soft_mask, enc_hw, sum_slot, wsum_slot, mask_slot=None):
    """One shot pooled straight INTO a bank slot: sum_slot [c], wsum_slot [1] and mask_slot [e] (nullable) are views of
    the bank's buffers and are accumulated in place (`+=`, as the reference's slot writes at :482-484)."""
    _need(feat, torch.float32, "feat")
    _need(soft_mask, torch.float32, "soft_mask")
    mh, mw = soft_mask.shape[-2:]
    eh, ew = enc_hw
    c = feat.shape[-1]
    lib = _lib.load()
    _lib.check(lib.nttt_fill_pool_accumulate(_ptr(feat), _ptr(soft_mask), mh, mw, eh, ew, c, _ptr(sum_slot),
                                             _ptr(wsum_slot), _ptr(mask_slot), _stream(feat.device)),
               "nttt_fill_pool_accumulate")


def fill_pool_batch(feats, soft_masks, enc_hw, sums, wsums, masks_lowres=None):
    """B reference shots in one launch: feats [B,E,C], soft_masks [B,S,S] -> sums [B,C], wsums [B], masks_lowres [B,E]
    (preallocated staging rows, plain stores)."""
    _need(feats, torch.float32, "feats")
    _need(soft_masks, torch.float32, "soft_masks")
    _need(sums, torch.float32, "sums")
    _need(wsums, torch.float32, "wsums")
    b, e, c = feats.shape
    mh, mw = soft_masks.shape[-2:]
    eh, ew = enc_hw
    assert e == eh * ew and soft_masks.shape[0] == b and sums.shape == (b, c) and wsums.shape == (b,)
    if masks_lowres is not None:
        _need(masks_lowres, torch.float32, "masks_lowres")
        assert masks_lowres.shape == (b, e)
    lib = _lib.load()
    _lib.check(lib.nttt_fill_pool_batch(_ptr(feats), _ptr(soft_masks), b, mh, mw, eh, ew, c, _ptr(sums), _ptr(wsums),
                                        _ptr(masks_lowres), _stream(feats.device)), "nttt_fill_pool_batch")


def fill_scatter(sums, wsums, masks_lowres, slot, feats_sum, mask_sum, masks=None):
    """Staged rows -> bank slots: feats_sum.view(-1,C)[slot[i]] += sums[i] etc. (slot < 0: skipped; unique slots)."""
    for t, name in ((sums, "sums"), (wsums, "wsums"), (feats_sum, "feats_sum"), (mask_sum, "mask_sum")):
        _need(t, torch.float32, name)
    _need(slot, torch.int32, "slot")
    n, c = sums.shape
    e = masks_lowres.shape[-1] if masks_lowres is not None else 1
    if masks is not None:
        _need(masks, torch.float32, "masks")
        _need(masks_lowres, torch.float32, "masks_lowres")
    else:
        masks_lowres = None
    lib = _lib.load()
    _lib.check(lib.nttt_fill_scatter(_ptr(sums), _ptr(wsums), _ptr(masks_lowres), _ptr(slot), n, c, e, _ptr(feats_sum),
                                     _ptr(mask_sum), _ptr(masks), _stream(sums.device)), "nttt_fill_scatter")


def fill_finalize(sums, wsum):
    _need(sums, torch.float32, "sums")
    _need(wsum, torch.float32, "wsum")
    n_cls, shots, c = sums.shape
    ins_avg = torch.empty_like(sums)
    avg = torch.empty((n_cls, c), dtype=torch.float32, device=sums.device)
    lib = _lib.load()
    _lib.check(lib.nttt_fill_finalize(_ptr(sums), _ptr(wsum), n_cls, shots, c, _ptr(ins_avg), _ptr(avg),
                                      _stream(sums.device)), "nttt_fill_finalize")
    return ins_avg, avg
