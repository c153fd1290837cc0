"""Candidate selection fused into the stage: best-of-3 plane choice per prompt
(`Sam2MatchingBaseline_noAMG.py:295-299`), the per-batch `cat` (:423-425) and `scores > iou_thr` (:428-431), consumed
in place from the decoder's per-batch tensors.  Checked against the oracle and against golden vectors produced by the
reference's real `_forward_sam` + `forward_test` (tests/golden/make_golden.py: MULTI_CASES)."""
import importlib
import types

import numpy as np
import pytest
import torch

from golden_util import MULTI_CASES, assert_rows_match, assert_same_ranking, load_multimask_case
from oracle import nttt_oracle as orc
from oracle import ref_torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def P():
    return importlib.import_module("no-time-to-train_b200")


@pytest.fixture(scope="module")
def ops(P):
    return importlib.import_module("no-time-to-train_b200.ops")


def _chunks(multi, bs):
    return [multi[i:i + bs].contiguous().to(DEV) for i in range(0, multi.shape[0], bs)]


def test_select_multimask_matches_argmax(ops, synth):
    multi, ious = synth.make_multimask_inputs(37, 4, seed=5)  # ragged: last chunk holds 5 prompts
    ious[3, 2] = float("nan")          # NaN counts as maximal
    ious[4, 1] = ious[4, 3] = float("nan")  # first NaN wins
    ious[5, 1:] = 0.25                 # all equal: first competing plane
    chunks = _chunks(multi, 8)
    mask_ptr, score = ops.select_multimask(ious.to(DEV), chunks, first=1)
    want = torch.argmax(ious[:, 1:], dim=-1) + 1
    plane_bytes = 256 * 256 * 4
    got_ptr = mask_ptr.cpu().tolist()
    for i in range(37):
        base = chunks[i // 8].data_ptr()
        assert got_ptr[i] == base + ((i % 8) * 4 + int(want[i])) * plane_bytes, i
    want_score = ious[torch.arange(37), want]
    assert np.array_equal(score.cpu().numpy(), want_score.numpy(), equal_nan=True)
    # first = 0 lets plane 0 compete
    mask_ptr0, score0 = ops.select_multimask(ious.to(DEV), chunks, first=0)
    want0 = torch.argmax(ious, dim=-1)
    assert np.array_equal(score0.cpu().numpy(), ious[torch.arange(37), want0].numpy(), equal_nan=True)


def test_threshold_pack_ptrs_with_gate_bit_exact(ops, synth):
    multi, ious = synth.make_multimask_inputs(24, 4, seed=6)
    multi[1, 2, 7, 9] = float("inf")
    chunks = _chunks(multi, 6)
    mask_ptr, score = ops.select_multimask(ious.to(DEV), chunks, first=1)
    thr = 0.6
    bits, area, box, stab, flags = ops.threshold_pack_ptrs(mask_ptr, (256, 256), gate=score, gate_min=thr)
    lr, sc, kept = orc.select_candidates(multi.numpy(), ious.numpy(), thr)
    assert 3 < len(kept) < 24
    mask, o_area, o_box, o_hi, o_lo = orc.threshold_stats(lr, 0.0, 1.0)
    b = bits.cpu().numpy().view(np.uint32)
    got = np.unpackbits(b.view(np.uint8), bitorder="little").reshape(24, 256, 256)
    assert np.array_equal(got[kept], mask)
    assert np.array_equal(area.cpu().numpy()[kept], o_area)
    assert np.array_equal(box.cpu().numpy()[kept].astype(np.int64), o_box)
    assert np.array_equal(stab.cpu().numpy()[kept, 0], o_hi) and np.array_equal(stab.cpu().numpy()[kept, 1], o_lo)
    dropped = np.setdiff1d(np.arange(24), kept)
    assert not got[dropped].any() and not area.cpu().numpy()[dropped].any() and not box.cpu().numpy()[dropped].any()


@pytest.mark.parametrize("name", MULTI_CASES)
@pytest.mark.parametrize("chunked", [True, False])
def test_stage_on_raw_decoder_output_matches_reference(P, name, chunked):
    """`nttt_match_image` with multi_ious + iou_thr on the RAW decoder output == the reference's `_forward_sam`
    (gather, cat, filter) followed by `forward_test`."""
    g, multi, ious, feat, cfg = load_multimask_case(name)
    stage = P.MatchingStage(DEV, P.StageConfig(nms_thr=0.5, num_out_instance=cfg["num_out_instance"], enc_hw=(37, 37)))
    stage.set_prototypes(feat.feats_ins_avg)
    lr_in = _chunks(multi, cfg["bs"]) if chunked else multi.to(DEV)
    out = stage.match(lr_in, None, feat.tar_feat.to(DEV), cfg["ori_hw"], iou_thr=cfg["iou_thr"],
                      multi_ious=ious.to(DEV), multi_first=1)
    assert_same_ranking(out["scores"].cpu().numpy(), out["labels"].cpu().numpy(), g["out_scores"], g["out_labels"],
                        what=name)
    assert_rows_match(dict(scores=out["scores"], labels=out["labels"], bboxes=out["bboxes"], masks=out["binary_masks"]),
                      dict(scores=g["out_scores"], labels=g["out_labels"], bboxes=g["out_bboxes"],
                           masks=g["out_masks_packed"]), what=name)
    # `index` refers to prompt numbers of the un-compacted grid
    _, _, kept = orc.select_candidates(multi.numpy(), ious.numpy(), cfg["iou_thr"])
    assert set(out["index"].cpu().tolist()) <= set(kept.tolist())


def test_multimask_graph_replay_matches_eager(P, synth):
    n, m, c = 64, 4, 128
    feat = synth.make_stage_inputs(8, c, 5, 2, (480, 640), seed=31, clustered=True)
    stage = P.MatchingStage(DEV, P.StageConfig(nms_thr=0.5, num_out_instance=10, enc_hw=(37, 37)))
    stage.set_prototypes(feat.feats_ins_avg)
    g = stage.graphed(n, c, (480, 640), iou_thr=0.5, n_multi=m).capture()
    for seed in (41, 42):
        multi, ious = synth.make_multimask_inputs(n, m, seed=seed)
        g.lr_masks.copy_(multi)
        g.multi_ious.copy_(ious)
        g.tar_feat.copy_(feat.tar_feat)
        out = g.replay().get()
        lr, sc, _ = ref_torch.select_candidates([multi], [ious], 0.5)
        eager = stage.match(lr.contiguous().to(DEV), sc.contiguous().to(DEV), feat.tar_feat.to(DEV), (480, 640),
                            low_latency=False)  # same accumulation order as the captured (throughput-mode) graph
        assert out["counts"] == eager["counts"]
        assert torch.equal(torch.nan_to_num(out["scores"]), torch.nan_to_num(eager["scores"]))
        assert torch.equal(out["labels"], eager["labels"])
        assert torch.equal(out["binary_masks"], eager["binary_masks"]) and torch.equal(out["bboxes"], eager["bboxes"])


@pytest.mark.parametrize("name", MULTI_CASES[:1])
def test_model_consumes_decoder_batches_in_place(P, name):
    """The model boundary with a stand-in SAM-2 predictor: `forward_test` hands the decoder's per-batch tensors to
    the stage without gathering / concatenating them, and gives the reference's result."""
    g, multi, ious, feat, cfg = load_multimask_case(name)
    bs = cfg["bs"]
    pps = int(round(multi.shape[0] ** 0.5))
    multi_d, ious_d = multi.to(DEV), ious.to(DEV)
    calls = {"i": 0, "ptrs": []}

    def decoder(**kw):
        assert kw["multimask_output"] is True and kw["output_all_masks"] is True
        i = calls["i"]
        calls["i"] += 1
        out = multi_d[i * bs:(i + 1) * bs].clone()  # a fresh tensor per batch, as the real decoder returns
        calls["ptrs"].append(out.data_ptr())
        return out, ious_d[i * bs:(i + 1) * bs].clone(), None, None

    prompt_encoder = lambda points, boxes, masks: (torch.zeros(bs, 2, 4, device=DEV), torch.zeros(bs, 4, 2, 2, device=DEV))
    prompt_encoder.get_dense_pe = lambda: torch.zeros(1, 4, 2, 2, device=DEV)
    pred = types.SimpleNamespace(
        forward_image=lambda imgs: {},
        _prepare_backbone_features=lambda bo: (None, [torch.zeros(16, 1, 4, device=DEV), torch.zeros(4, 1, 4, device=DEV)],
                                                None, [(4, 4), (2, 2)]),
        sam_prompt_encoder=prompt_encoder, sam_mask_decoder=decoder)

    class Model(P.Sam2MatchingBaselineNoAMG):
        def _extract_target_features(self, tar_img, device):
            return feat.tar_feat.to(device), tar_img.to(device)

    c = feat.tar_feat.shape[1]
    m = Model(sam2_infer_cfgs=dict(points_per_side=pps, testing_point_bs=bs, iou_thr=cfg["iou_thr"], nms_thr=0.5,
                                   num_out_instance=cfg["num_out_instance"], kmeans_k=2, n_pca_components=2,
                                   cls_num_per_mask=1),
              memory_bank_cfg=dict(enable=True, category_num=cfg["n_cls"], length=feat.feats_ins_avg.shape[1]),
              encoder_geometry=(518, 14, c), predictor=pred, device=DEV)
    m.memory_bank.feats_ins_avg.copy_(feat.feats_ins_avg)
    m.memory_bank.postprocessed[0] = True
    info = dict(ori_height=cfg["ori_hw"][0], ori_width=cfg["ori_hw"][1], file_name="x", id=0)
    out = m([dict(data_mode="test", target_img=torch.zeros(3, 64, 64), target_img_info=info)])[0]
    assert calls["i"] == multi.shape[0] // bs
    assert_same_ranking(out["scores"].cpu().numpy(), out["labels"].cpu().numpy(), g["out_scores"], g["out_labels"],
                        what="model")
    assert_rows_match(dict(scores=out["scores"], labels=out["labels"], bboxes=out["bboxes"], masks=out["binary_masks"]),
                      dict(scores=g["out_scores"], labels=g["out_labels"], bboxes=g["out_bboxes"],
                           masks=g["out_masks_packed"]), what="model")
    # the reference-compatible seam (compacted triple) gives the same candidates
    calls["i"] = 0
    lr, sc, _ = m._forward_sam(torch.zeros(1, 3, 64, 64, device=DEV))
    assert np.array_equal(sc.cpu().numpy(), g["sel_pred_ious"]) and lr.shape[0] == int(g["sel_count"])
