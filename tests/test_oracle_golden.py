"""CPU tests: both oracles (torch restatement, C/numpy restatement) against vectors produced by the
REAL reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from golden_util import (assert_rows_match, FILL_CASES, MULTI_CASES, STAGE_CASES, assert_close_rel, assert_same_ranking, load_case,
                         load_multimask_case, sha_bool, sha_f32, GOLDEN_DIR)
from oracle import nttt_oracle as orc
from oracle import ref_torch

import importlib
import os


def _check_against_golden(g, out, aux, what):
    assert_close_rel(aux["sim"], g["sim"], what=what + " sim")
    assert_close_rel(aux["obj_feats"], g["obj_feats"], what=what + " obj_feats")
    assert np.array_equal(np.asarray(aux["lr_boxes"]), g["lr_boxes"]), what + " lr boxes"
    out_num = min(8 * int(g["spec"][8]), int(g["spec"][0]))
    assert np.array_equal(np.asarray(aux["keep"]), g["nms_keep_full"][:out_num]), what + " nms keep"
    if "ios" in g:
        assert np.array_equal(np.asarray(aux["full_area"]), g["full_area"]), what + " full-res areas"
        assert np.array_equal(np.asarray(aux["full_boxes"]), g["full_boxes"]), what + " full-res boxes"
        assert_close_rel(aux["ios"], g["ios"], what=what + " ios")
    assert_same_ranking(out["scores"], out["labels"], g["out_scores"], g["out_labels"], what=what)
    assert_rows_match(dict(scores=out["scores"], labels=out["labels"], bboxes=out["bboxes"],
                           masks=np.asarray(out["binary_masks"]).astype(bool)),
                      dict(scores=g["out_scores"], labels=g["out_labels"], bboxes=g["out_bboxes"],
                           masks=g["out_masks_packed"]), what=what)


@pytest.mark.parametrize("name", STAGE_CASES)
def test_torch_restatement_matches_reference(name):
    g, inp, cfg = load_case(name)
    with torch.inference_mode():
        out = ref_torch.match_image(inp.lr_masks, inp.pred_ious, inp.tar_feat, inp.feats_ins_avg,
                                    ref_torch.StageConfig(num_out_instance=cfg["num_out_instance"]), inp.ori_hw)
        stab = ref_torch.stability_score(inp.lr_masks, 0.0, 1.0)
    aux = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out["aux"].items()}
    res = {k: v.numpy() for k, v in out.items() if k != "aux"}
    _check_against_golden(g, res, aux, name + " [torch]")
    assert np.array_equal(stab.numpy(), g["stability"], equal_nan=True)
    if "full_masks_sha" in g:
        full = ref_torch.upsample_threshold(inp.lr_masks[out["aux"]["sel_index"]], inp.ori_hw)
        assert sha_bool(full) == str(g["full_masks_sha"])


@pytest.mark.parametrize("name", STAGE_CASES)
def test_c_oracle_matches_reference(name):
    g, inp, cfg = load_case(name)
    res = orc.match_image(inp.lr_masks.numpy(), inp.pred_ious.numpy(), inp.tar_feat.numpy(),
                          inp.feats_ins_avg.numpy(), inp.ori_hw, num_out_instance=cfg["num_out_instance"])
    _check_against_golden(g, res, res, name + " [C]")
    with np.errstate(invalid="ignore", divide="ignore"):
        stab = res["stab_hi"].astype(np.float32) / res["stab_lo"].astype(np.float32)
    assert np.array_equal(stab, g["stability"], equal_nan=True)
    if "full_masks_sha" in g:
        # bit-exact thresholded full-resolution masks for EVERY selected mask, not only the outputs
        full = orc.aa_resize_threshold(inp.lr_masks.numpy()[res["sel_index"]], inp.ori_hw)
        assert sha_bool(full) == str(g["full_masks_sha"])
        # integer intersections == fp32 matmul counts of the reference (checked through ios above), and
        # symmetric
        assert np.array_equal(res["inter"], res["inter"].T)


@pytest.mark.parametrize("out_hw", [(1024, 1024), (480, 640), (427, 640), (333, 500), (200, 180), (256, 256), (2048, 1536)])
def test_aa_resize_bit_exact_vs_aten(out_hw):
    """The C recipe reproduces aten's `_upsample_bilinear2d_aa` bit-for-bit (float outputs, not only signs)
    for every scale <= 2 per axis (all up-scaling and mild down-scaling)."""
    gen = torch.Generator().manual_seed(7)
    src = torch.randn(3, 256, 256, generator=gen) * 5
    ref = torch.nn.functional.interpolate(src[:, None], size=out_hw, mode="bilinear", align_corners=False,
                                          antialias=True)[:, 0]
    got = orc.aa_resize(src.numpy(), out_hw)
    assert np.array_equal(got.view(np.uint32), ref.numpy().view(np.uint32))


@pytest.mark.parametrize("out_hw", [(100, 700), (100, 100), (120, 256), (64, 64)])
def test_aa_resize_heavy_downscale_masks_equal(out_hw):
    """Down-scaling by more than 2x: aten's CPU kernel sums its >4 taps in a different association than
    the CUDA kernel whose recipe the oracle follows, so float outputs differ by an ulp; the thresholded
    masks (the contract) are still identical."""
    gen = torch.Generator().manual_seed(7)
    src = torch.randn(3, 256, 256, generator=gen) * 5
    ref = torch.nn.functional.interpolate(src[:, None], size=out_hw, mode="bilinear", align_corners=False,
                                          antialias=True)[:, 0]
    got = orc.aa_resize_threshold(src.numpy(), out_hw)
    assert np.array_equal(got.astype(bool), (ref > 0).numpy())
    assert np.allclose(orc.aa_resize(src.numpy(), out_hw), ref.numpy(), rtol=0, atol=1e-5)


def test_aa_feature_upsample_37_to_256_bit_exact():
    gen = torch.Generator().manual_seed(8)
    src = torch.randn(4, 37, 37, generator=gen)
    ref = torch.nn.functional.interpolate(src[None], size=(256, 256), mode="bilinear", align_corners=False,
                                          antialias=True)[0]
    got = orc.aa_resize(src.numpy(), (256, 256))
    assert np.array_equal(got.view(np.uint32), ref.numpy().view(np.uint32))


def test_nms_semantics_probed_in_survey():
    """Tie / threshold / zero-area semantics of torchvision nms."""
    from torchvision.ops import nms
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [0, 0, 10, 10], [50, 50, 60, 60],
                      [0, 0, 10, 20], [7, 7, 7, 30], [0, 0, 0, 0]], np.float32)
    scores = np.array([.9, .9, .9, .9, .5, .4, .3], np.float32)
    labels = np.zeros(7, np.int64)
    for thr in (0.5, 0.3, 0.7):
        want = nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
        got = orc.box_nms(boxes, scores, labels, thr)
        assert np.array_equal(got, want)
    # IoU exactly == threshold is NOT suppressed: boxes 0 and 4 have IoU 100/200 = 0.5
    assert 4 in orc.box_nms(boxes, scores, labels, 0.5)


def test_nms_random_vs_torchvision():
    from torchvision.ops import batched_nms
    gen = torch.Generator().manual_seed(3)
    for n in (1, 17, 300, 999):
        xy = torch.randint(0, 200, (n, 2), generator=gen)
        wh = torch.randint(0, 56, (n, 2), generator=gen)
        boxes = torch.cat([xy, xy + wh], 1).float()
        scores = torch.rand(n, generator=gen)
        labels = torch.randint(0, 4, (n,), generator=gen)
        want = batched_nms(boxes, scores, labels, 0.5).numpy()
        got = orc.box_nms(boxes.numpy(), scores.numpy(), labels.numpy(), 0.5)
        assert np.array_equal(got, want), n


@pytest.mark.parametrize("name", FILL_CASES)
def test_bank_restatement_matches_reference(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    n_cls, shots, c, seed, e_side, img_side = g["spec"].tolist()
    synth = importlib.import_module("no-time-to-train_b200.synth")
    feats, _ = synth.make_ref_shots(n_cls, shots, e_side * e_side, c, seed=seed)
    bank = ref_torch.RawBank(n_cls, shots, e_side * e_side, c)
    for (ci, li), soft in zip(g["order"].tolist(), g["soft_masks"]):
        m = torch.nn.functional.interpolate(torch.from_numpy(soft)[None, None], size=(e_side, e_side),
                                            mode="nearest").reshape(1, -1)
        ref_torch.bank_fill(bank, [ci], feats[ci, li][None], m)
    assert np.array_equal(bank.fill_counts.numpy(), g["fill_counts"])
    assert np.array_equal(bank.masks.numpy(), g["masks_lowres"])
    feats_avg, feats_ins_avg = ref_torch.bank_postprocess(bank)
    assert_close_rel(feats_avg.numpy(), g["feats_avg"], rtol=1e-6, what="feats_avg")
    assert_close_rel(feats_ins_avg.numpy(), g["feats_ins_avg"], rtol=1e-6, what="feats_ins_avg")


@pytest.mark.parametrize("name", MULTI_CASES)
def test_candidate_selection_restatements_match_reference(name):
    """Best-of-3 plane choice + cat + `> iou_thr` (the reference's real `_forward_sam`), then the stage."""
    g, multi, ious, feat, cfg = load_multimask_case(name)
    bs = cfg["bs"]
    chunks = [multi[i:i + bs] for i in range(0, multi.shape[0], bs)]
    iou_chunks = [ious[i:i + bs] for i in range(0, multi.shape[0], bs)]
    lr, sc, kept = ref_torch.select_candidates(chunks, iou_chunks, cfg["iou_thr"])
    assert lr.shape[0] == int(g["sel_count"])
    assert np.array_equal(sc.numpy(), g["sel_pred_ious"])
    assert sha_f32(lr) == str(g["sel_masks_sha"])
    lr_c, sc_c, kept_c = orc.select_candidates(multi.numpy(), ious.numpy(), cfg["iou_thr"])
    assert np.array_equal(kept_c, kept.numpy()) and np.array_equal(sc_c, g["sel_pred_ious"])
    assert sha_f32(lr_c) == str(g["sel_masks_sha"])
    with torch.inference_mode():
        out = ref_torch.match_image(lr, sc, feat.tar_feat, feat.feats_ins_avg,
                                    ref_torch.StageConfig(num_out_instance=cfg["num_out_instance"]), cfg["ori_hw"])
    assert_same_ranking(out["scores"].numpy(), out["labels"].numpy(), g["out_scores"], g["out_labels"], what=name)
    masks = out["binary_masks"].numpy().astype(np.uint8)
    assert np.array_equal(np.packbits(masks.reshape(masks.shape[0], -1), axis=-1), g["out_masks_packed"])
    assert np.array_equal(out["bboxes"].numpy(), g["out_bboxes"])


def test_candidate_selection_nan_and_ties():
    """torch.argmax semantics the kernel must share: first maximal value wins, NaN counts as maximal."""
    ious = torch.tensor([[0.9, 0.5, 0.5, 0.5], [0.1, 0.2, float("nan"), 0.9], [0.0, float("nan"), float("nan"), 1.0],
                         [0.3, 0.1, 0.7, 0.7]])
    multi = torch.arange(4 * 4 * 4, dtype=torch.float32).reshape(4, 4, 2, 2)
    want = torch.argmax(ious[:, 1:], dim=-1) + 1
    assert want.tolist() == [1, 2, 1, 2]
    lr_c, sc_c, kept_c = orc.select_candidates(multi.numpy(), ious.numpy(), 0.4)
    lr_t, sc_t, kept_t = ref_torch.select_candidates([multi], [ious], 0.4)
    assert np.array_equal(kept_c, kept_t.numpy()) and kept_c.tolist() == [0, 3]  # NaN > thr is False
    assert np.array_equal(lr_c, lr_t.numpy())
