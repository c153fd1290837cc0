"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-golden (GPU) tests."""
import hashlib
import importlib
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STAGE_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("stage_") and f.endswith(".npz"))
FILL_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("fill_") and f.endswith(".npz"))

MULTI_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("stagemm_") and f.endswith(".npz"))

RTOL = 1e-3  # BASELINE.json north_star: pooled features / prototypes / similarities within 1e-3 relative


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    n, c, n_cls, shots, oh, ow, seed, degenerate, num_out, clustered = g["spec"].tolist()
    synth = importlib.import_module("no-time-to-train_b200.synth")
    inp = synth.make_stage_inputs(n, c, n_cls, shots, (oh, ow), seed=seed, clustered=bool(clustered),
                                  degenerate=int(degenerate))
    h = hashlib.sha256()
    for t in (inp.lr_masks, inp.pred_ious, inp.tar_feat, inp.feats_ins_avg):
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    assert h.hexdigest() == str(g["inputs_sha"]), "synthetic generator drifted from the golden inputs"
    return g, inp, dict(num_out_instance=num_out, n_cls=n_cls)


def load_multimask_case(name):
    """Candidate-selection golden (real `_forward_sam` + `forward_test` over synthetic RAW decoder output)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    pps, bs, m, c, n_cls, shots, oh, ow, seed, num_out = g["spec"].tolist()
    synth = importlib.import_module("no-time-to-train_b200.synth")
    multi, ious = synth.make_multimask_inputs(pps * pps, m, seed=seed)
    feat = synth.make_stage_inputs(8, c, n_cls, shots, (oh, ow), seed=seed + 1, clustered=True)
    h = hashlib.sha256()
    for t in (multi, ious, feat.tar_feat, feat.feats_ins_avg):
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    assert h.hexdigest() == str(g["inputs_sha"]), "synthetic generator drifted from the golden inputs"
    cfg = dict(bs=bs, iou_thr=float(g["iou_thr"]), num_out_instance=num_out, n_cls=n_cls, ori_hw=(oh, ow))
    return g, multi, ious, feat, cfg


def sha_f32(t) -> str:
    a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    return hashlib.sha256(np.ascontiguousarray(a.astype(np.float32)).tobytes()).hexdigest()


def sha_bool(t) -> str:
    a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    return hashlib.sha256(np.ascontiguousarray(a.astype(np.bool_)).tobytes()).hexdigest()


def assert_close_rel(actual, expected, rtol=RTOL, what=""):
    """|a-e| <= rtol * max|e| row-wise scale (vectors are unit-norm or cosine-valued), NaN == NaN."""
    a = np.asarray(actual, dtype=np.float64)
    e = np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    assert np.array_equal(np.isnan(a), np.isnan(e)), f"{what}: NaN pattern differs"
    scale = max(np.nanmax(np.abs(e)), 1e-30) if e.size else 1.0
    err = np.nanmax(np.abs(a - e)) / scale if e.size else 0.0
    assert err <= rtol, f"{what}: max relative error {err:.3e} > {rtol}"


def assert_same_ranking(scores_a, labels_a, scores_e, labels_e, rtol=RTOL, what=""):
    """Final per-image rankings must be identical; permutations are tolerated only inside groups whose
    reference scores tie within the float tolerance (argsort is unstable in the reference, :674-675)."""
    sa, se = np.asarray(scores_a, np.float64), np.asarray(scores_e, np.float64)
    la, le = np.asarray(labels_a), np.asarray(labels_e)
    assert sa.shape == se.shape, f"{what}: K_out {sa.shape} vs {se.shape}"
    assert_close_rel(sa, se, rtol, what + " scores")
    if np.array_equal(la, le):
        return
    bad = np.nonzero(la != le)[0]
    for i in bad:
        near = np.abs(se - se[i]) <= rtol * max(abs(se[i]), 1e-6)
        assert le[i] in la[near] and la[i] in le[near], f"{what}: label order differs at rank {i} outside a tie group"


def _packed_rows(masks):
    """[K, H, W] bool/uint8 (numpy or torch) -> [K, ceil(HW/8)] uint8 rows, the layout the golden files store."""
    m = masks.detach().cpu().numpy() if isinstance(masks, torch.Tensor) else np.asarray(masks)
    if m.dtype == np.uint8 and m.ndim == 2:
        return m  # already packed
    return np.packbits(m.reshape(m.shape[0], -1).astype(np.uint8), axis=-1)


def assert_rows_match(got, ref, rtol=RTOL, what=""):
    """ALWAYS compares labels, boxes and masks of every output row with the reference — unconditionally.

    `got` / `ref`: dicts with `scores [K]`, `labels [K]`, `bboxes [K,4]`, `masks` ([K,H,W] bool or packed rows) and
    optionally `index [K]`.  The reference ranks with an unstable `argsort` (`Sam2MatchingBaseline_noAMG.py:674-675`)
    and our float scores differ from its in the last bits, so rows whose reference scores tie within the tolerance
    (and NaN rows among themselves) may be permuted: every output row must find an unused reference row INSIDE its
    tie group with the same label, box, mask (and index).  Returns and prints the number of permuted rows."""
    def arr(x):
        return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    sa, se = arr(got["scores"]).astype(np.float64), arr(ref["scores"]).astype(np.float64)
    la, le = arr(got["labels"]), arr(ref["labels"])
    ba, be = arr(got["bboxes"]).astype(np.int64), arr(ref["bboxes"]).astype(np.int64)
    ma, me = _packed_rows(got["masks"]), _packed_rows(ref["masks"])
    ia = arr(got["index"]) if "index" in got and "index" in ref else None
    ie = arr(ref["index"]) if ia is not None else None
    k = sa.shape[0]
    assert se.shape[0] == k and la.shape[0] == k and le.shape[0] == k, f"{what}: K_out {sa.shape} vs {se.shape}"
    assert ba.shape == be.shape and ma.shape == me.shape, f"{what}: box/mask shapes {ba.shape}/{ma.shape} vs {be.shape}/{me.shape}"
    assert_close_rel(sa, se, rtol, what + " scores")
    scale = max(np.nanmax(np.abs(se)), 1e-30) if k and not np.all(np.isnan(se)) else 1.0
    used = np.zeros(k, dtype=bool)
    moved = 0
    for i in range(k):
        if np.isnan(se[i]):
            group = np.isnan(se)
        else:
            with np.errstate(invalid="ignore"):
                group = np.abs(se - se[i]) <= 2 * rtol * scale
        cand = sorted(np.nonzero(group & ~used)[0].tolist(), key=lambda j: (j != i, abs(j - i)))
        for j in cand:
            if la[i] == le[j] and np.array_equal(ba[i], be[j]) and np.array_equal(ma[i], me[j]) and \
                    (ia is None or ia[i] == ie[j]):
                used[j] = True
                moved += int(j != i)
                break
        else:
            raise AssertionError(f"{what}: output row {i} (label {la[i]}, box {ba[i].tolist()}) has no identical "
                                 f"counterpart among the {len(cand)} unused reference rows of its tie group")
    print(f"[{what}] {k} output rows compared (labels, boxes, masks), {moved} inside-tie permutations")
    return moved
