"""Result encoding after the stage.

Mirror of the reference's `_output_inqueue` + `encode_results`
(`no_time_to_train/pl_wrapper/sam2matcher_pl.py:144-158`, `no_time_to_train/dataset/coco_ref_dataset.py:590-613`):
one COCO result dict per output instance.  The reference copies the dense `[K_out,H,W]` bool masks to the host
(105 MB per image) and runs pycocotools per mask; here the `segmentation` strings were already produced on the
device by `nttt_rle_encode` (fused into `nttt_match_image`), so the host reads a few KB per image.
"""
from __future__ import annotations

from .matching import PendingResult


def box_xyxy_to_xywh(box):
    """`_box_xyxy_to_xywh` (`coco_ref_dataset.py:116-128`): width/height are plain differences (no +1)."""
    x1, y1, x2, y2 = box
    return [x1, y1, x2 - x1, y2 - y1]


def rle_encode_host(mask) -> dict:
    """Host mirror of pycocotools `encode(np.asfortranarray(mask))` for callers that hold a dense mask and did not ask
    the stage for RLE output (`coco_ref_dataset.py:601-604` does this on the CPU for every mask): column-major runs
    starting with zeros, then rleToString's 5-bit groups (chars 48..111).  The fused device encoder
    (`nttt_rle_encode`) produces the same strings without the dense D2H copy."""
    import numpy as np
    m = np.asarray(mask.cpu() if hasattr(mask, "cpu") else mask).astype(np.uint8)
    h, w = m.shape
    flat = m.reshape(-1, order="F")
    edges = np.flatnonzero(np.diff(np.concatenate([[0], flat]).astype(np.int8))) if flat.size else np.zeros(0, np.int64)
    counts = np.diff(np.concatenate([[0], edges, [flat.size]])).astype(np.int64)
    chars = bytearray()
    for i, c in enumerate(counts.tolist()):
        x = c - (counts[i - 2] if i > 2 else 0)
        x = int(x)
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            chars.append(ch + 48)
    return dict(size=[int(h), int(w)], counts=chars.decode("ascii"))


def encode_results(pending: PendingResult, img_id, cat_inds_to_ids=None) -> list:
    """-> [{"image_id", "category_id", "bbox" (xywh), "score", "segmentation": {"size", "counts"}}, ...] in the
    stage's output order.  `cat_inds_to_ids` maps label indices to dataset category ids (identity if None)."""
    out = pending.get()
    segs = pending.rle_segmentations()
    img_id = int(img_id) if str(img_id).isdigit() else img_id  # (:594)
    scores = out["scores"].cpu().tolist()
    labels = out["labels"].cpu().tolist()
    boxes = out["bboxes"].cpu().tolist()
    results = []
    for score, label, box, seg in zip(scores, labels, boxes, segs):
        cat = int(label) if cat_inds_to_ids is None else int(cat_inds_to_ids[int(label)])
        results.append(dict(image_id=img_id, category_id=cat, bbox=box_xyxy_to_xywh(box), score=float(score),
                            segmentation=seg))
    return results
