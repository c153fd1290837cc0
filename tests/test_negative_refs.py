"""Negative-reference scoring: `with_negative_refs=True` path of the reference
(`compute_sim_global_avg_with_neg`, `matching_baseline_utils.py:906-941`; modes `fill_memory_neg`, `test`,
`test_support` of `Sam2MatchingBaselineNoAMG.forward`, `:728-763`).

Golden vectors come from the real reference's `forward_test(with_negative=True)` (tests/golden/make_golden.py).
"""
import hashlib
import importlib
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, assert_close_rel, assert_rows_match, assert_same_ranking
from oracle import ref_torch

NEG_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("stageneg_") and f.endswith(".npz"))
DEV = "cuda:0"


def load_neg_case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    n, c, n_cls, l_neg, oh, ow, seed, num_out = g["spec"].tolist()[:8]
    degenerate = int(g["spec"][8]) if g["spec"].shape[0] > 8 else 0  # no zero guard in the reference: NaN rows
    synth = importlib.import_module("no-time-to-train_b200.synth")
    inp = synth.make_stage_inputs(n, c, n_cls, 2, (oh, ow), seed=seed, clustered=True, degenerate=degenerate)
    h = hashlib.sha256()
    for t in (inp.lr_masks, inp.pred_ious, inp.tar_feat, inp.feats_ins_avg):
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    assert h.hexdigest() == str(g["inputs_sha"])
    return g, inp, num_out


@pytest.mark.parametrize("name", NEG_CASES)
def test_torch_port_negative_matches_reference(name):
    g, inp, num_out = load_neg_case(name)
    with torch.inference_mode():
        out = ref_torch.match_image(inp.lr_masks, inp.pred_ious, inp.tar_feat, inp.feats_ins_avg,
                                    ref_torch.StageConfig(num_out_instance=num_out), inp.ori_hw,
                                    negative=dict(feats_avg=torch.from_numpy(g["feats_avg"]),
                                                  feats_ins_avg_neg=torch.from_numpy(g["feats_ins_avg_neg"])))
    assert_close_rel(out["aux"]["sim"].numpy(), g["sim"], what="sim")
    assert_rows_match(dict(scores=out["scores"], labels=out["labels"], bboxes=out["bboxes"], masks=out["binary_masks"]),
                      dict(scores=g["out_scores"], labels=g["out_labels"], bboxes=g["out_bboxes"],
                           masks=g["out_masks_packed"]), what=name + " (torch port)")
    if "nms_keep_full" in g:  # degenerate cases: the NaN row's label and its place in the NMS keep list
        assert np.array_equal(out["aux"]["keep"].numpy(), g["nms_keep_full"][:out["aux"]["keep"].numel()])
        assert np.array_equal(out["aux"]["labels_all"].numpy(), g["labels_all"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NEG_CASES)
def test_pipeline_negative_matches_reference(name):
    P = importlib.import_module("no-time-to-train_b200")
    g, inp, num_out = load_neg_case(name)
    stage = P.MatchingStage(DEV, P.StageConfig(nms_thr=0.5, num_out_instance=num_out, enc_hw=(37, 37)))
    stage.set_prototypes_with_negatives(torch.from_numpy(g["feats_avg"]), torch.from_numpy(g["feats_ins_avg_neg"]))
    out = stage.match(inp.lr_masks.to(DEV), inp.pred_ious.to(DEV), inp.tar_feat.to(DEV), inp.ori_hw, taps=True)
    assert_close_rel(out["taps"]["sim"].cpu().numpy(), g["sim"], what="sim")
    assert_rows_match(dict(scores=out["scores"], labels=out["labels"], bboxes=out["bboxes"], masks=out["binary_masks"]),
                      dict(scores=g["out_scores"], labels=g["out_labels"], bboxes=g["out_bboxes"],
                           masks=g["out_masks_packed"]), what=name)
    if "nms_keep_full" in g:
        # degenerate inputs: the reference's negative path has no zero guard (matching_baseline_utils.py:925), so the
        # empty low-res mask has a NaN feature row, NaN similarities, a top-k over an all-NaN row and a NaN score that
        # `> 0` drops.  torch leaves the label of an all-NaN row unspecified (the reference got 0 with 80 classes and 28
        # with 40); it cannot matter: the empty mask's box is [0,0,0,0], whose intersection with any box is 0, so it
        # neither suppresses nor is suppressed under any label — it only keeps its slot in the NMS keep list.
        sim = out["taps"]["sim"].cpu().numpy()
        assert np.isnan(sim[0]).all() and np.isnan(g["sim"][0]).all()
        assert np.isnan(out["taps"]["obj_feats"].cpu().numpy()[0]).all()
        assert out["counts"]["n_keep"] == min(len(g["nms_keep_full"]), 8 * num_out)
        assert 0 in g["nms_keep_full"].tolist()
        lab = np.where(np.isnan(sim).all(1), 0, np.nanargmax(np.where(np.isnan(sim), -np.inf, sim), axis=1))
        for i in np.nonzero(lab != g["labels_all"])[0]:  # only float near-ties may flip a label
            if i == 0:
                continue
            assert abs(g["sim"][i, lab[i]] - g["sim"][i, g["labels_all"][i]]) <= 1e-5
    # switching back to positive-only prototypes must disable the negative term
    stage.set_prototypes(inp.feats_ins_avg)
    ref = ref_torch.match_image(inp.lr_masks, inp.pred_ious, inp.tar_feat, inp.feats_ins_avg,
                                ref_torch.StageConfig(num_out_instance=num_out), inp.ori_hw)
    out2 = stage.match(inp.lr_masks.to(DEV), inp.pred_ious.to(DEV), inp.tar_feat.to(DEV), inp.ori_hw, taps=True)
    assert_close_rel(out2["taps"]["sim"].cpu().numpy(), ref["aux"]["sim"].numpy(), what="sim (positive only)")


@pytest.mark.gpu
def test_similarity_neg_top1_entry():
    """The stand-alone C entry against the torch formula."""
    ops = importlib.import_module("no-time-to-train_b200.ops")
    gen = torch.Generator().manual_seed(3)
    n, c, n_cls, l_neg = 130, 384, 7, 3
    obj = torch.nn.functional.normalize(torch.randn(n, c, generator=gen), dim=-1)
    pos = torch.nn.functional.normalize(torch.randn(n_cls, c, generator=gen), dim=-1)
    neg = torch.nn.functional.normalize(torch.randn(n_cls * l_neg, c, generator=gen), dim=-1)
    sim, top_score, top_label = ops.similarity_neg_top1(obj.to(DEV), pos.to(DEV), neg.to(DEV), l_neg, 0.8)
    sp = (obj @ pos.t()).clamp(min=0)
    sn = (obj @ neg.t()).clamp(min=0).reshape(n, n_cls, l_neg).max(-1).values
    want = sp * torch.exp(-1.0 * (sn - sp).clamp(min=0) / 0.8)
    assert_close_rel(sim.cpu().numpy(), want.numpy(), what="sim_neg")
    assert np.array_equal(top_label.cpu().numpy(), sim.cpu().numpy().argmax(1))


@pytest.mark.gpu
def test_top1_nan_semantics_follow_torch():
    """torch.topk treats NaN as the maximum and torch.clamp / torch.max propagate it.  Which of several NaNs (or of
    several equal values) torch returns is unspecified — its CPU kernel gave column 17 or 31 row by row here — so the
    contract is: the label is a NaN column whenever the row has one (ours: the lowest such column)."""
    ops = importlib.import_module("no-time-to-train_b200.ops")
    gen = torch.Generator().manual_seed(4)
    n, c, n_cls, l_neg = 70, 128, 40, 2
    obj = torch.nn.functional.normalize(torch.randn(n, c, generator=gen), dim=-1)
    pos = torch.nn.functional.normalize(torch.randn(n_cls, c, generator=gen), dim=-1)
    neg = torch.nn.functional.normalize(torch.randn(n_cls * l_neg, c, generator=gen), dim=-1)
    obj[5] = float("nan")                 # a NaN feature row: every similarity NaN -> label 0, score NaN
    pos_nan = pos.clone()
    pos_nan[17] = float("nan")            # a NaN prototype: column 17 is NaN in every row -> label 17 everywhere
    pos_nan[31] = float("nan")
    for fn, args in ((ops.similarity_top1, (obj.to(DEV), pos_nan.to(DEV))),
                     (ops.similarity_neg_top1, (obj.to(DEV), pos_nan.to(DEV), neg.to(DEV), l_neg, 0.8))):
        sim, top_score, top_label = fn(*args)
        sim = sim.cpu()
        assert torch.isnan(sim[5]).all() and torch.isnan(sim[:, 17]).all() and torch.isnan(sim[:, 31]).all()
        assert not torch.isnan(sim[6, :17]).any()
        want_v, want_i = torch.topk(sim, k=1)
        assert torch.isnan(top_score.cpu()).all() and torch.isnan(want_v).all()
        lab = top_label.cpu()
        assert int(lab[5]) == 0 and (lab[torch.arange(n) != 5] == 17).all()
        rows = torch.arange(n) != 5
        assert set(want_i.flatten()[rows].tolist()) <= {17, 31}, "torch.topk itself must pick a NaN column"
    # a NaN only on the negative side propagates through max / clamp / exp as in torch
    neg_nan = neg.clone()
    neg_nan[2 * l_neg + 1] = float("nan")  # class 2, second negative slot
    sim, top_score, top_label = ops.similarity_neg_top1(obj.to(DEV), pos.to(DEV), neg_nan.to(DEV), l_neg, 0.8)
    sp = (obj @ pos.t()).clamp(min=0)
    sn = (obj @ neg_nan.t()).clamp(min=0).reshape(n, n_cls, l_neg).max(-1).values
    want = sp * torch.exp(-1.0 * (sn - sp).clamp(min=0) / 0.8)
    assert_close_rel(sim.cpu().numpy(), want.numpy(), what="sim_neg with NaN negatives")
    rows = torch.arange(n) != 5  # row 5 is NaN in every column: its label is the lowest index, 0
    assert torch.isnan(sim.cpu()[:, 2]).all() and torch.isnan(top_score.cpu()).all()
    assert (top_label.cpu()[rows] == 2).all() and int(top_label[5]) == 0


@pytest.mark.gpu
def test_model_negative_modes():
    """fill_memory / fill_memory_neg / test_support / test through the reference's forward() boundary."""
    P = importlib.import_module("no-time-to-train_b200")
    synth = P.synth
    n_cls, shots, l_neg, c = 3, 2, 2, 64
    feats, masks = synth.make_ref_shots(n_cls, shots + l_neg, 1369, c, seed=19)
    inp = synth.make_stage_inputs(n=32, c=c, n_cls=n_cls, shots=shots, ori_hw=(120, 160), seed=20)

    class Model(P.Sam2MatchingBaselineNoAMG):
        def _forward_encoder(self, imgs):
            ci, li = self._next
            return feats[ci, li].to(imgs.device)[None]

        def _extract_target_features(self, tar_img, device):
            return inp.tar_feat.to(device), tar_img.to(device)

        def _forward_sam(self, imgs):
            return inp.lr_masks.to(imgs.device), inp.pred_ious.to(imgs.device), None

    m = Model(sam2_infer_cfgs=dict(points_per_side=32, testing_point_bs=256, iou_thr=0.4, nms_thr=0.5,
                                   num_out_instance=10, kmeans_k=2, n_pca_components=2, cls_num_per_mask=1,
                                   with_negative_refs=True),
              memory_bank_cfg=dict(enable=True, category_num=n_cls, length=shots, length_negative=l_neg),
              encoder_geometry=(518, 14, c), device=DEV)
    info = dict(ori_height=120, ori_width=160, file_name="x", id=0)

    def test_dict(mode):
        return [dict(data_mode=mode, target_img=torch.zeros(3, 32, 32), target_img_info=info)]

    for li in range(shots):
        for ci in range(n_cls):
            m._next = (ci, li)
            m([dict(data_mode="fill_memory", refs_by_cat={ci: dict(imgs=torch.rand(1, 3, 64, 64),
                                                                   masks=masks[ci, li].reshape(1, 37, 37))})])
    m.postprocess_memory()
    with pytest.raises(RuntimeError, match="Negative memory is not ready"):
        m(test_dict("test"))
    # test_support: positives ready, negatives not post-processed -> positive-only scoring
    sup = m(test_dict("test_support"))[0]
    raw = ref_torch.RawBank(n_cls, shots, 1369, c)
    for li in range(shots):
        for ci in range(n_cls):
            ref_torch.bank_fill(raw, [ci], feats[ci, li][None], masks[ci, li][None])
    pos_avg, pos_ins = ref_torch.bank_postprocess(raw)
    ref_sup = ref_torch.match_image(inp.lr_masks, inp.pred_ious, inp.tar_feat, pos_ins,
                                    ref_torch.StageConfig(num_out_instance=10), (120, 160))
    assert_same_ranking(sup["scores"].cpu().numpy(), sup["labels"].cpu().numpy(), ref_sup["scores"].numpy(),
                        ref_sup["labels"].numpy(), what="test_support")
    for li in range(l_neg):
        for ci in range(n_cls):
            m._next = (ci, shots + li)
            m([dict(data_mode="fill_memory_neg", refs_by_cat={ci: dict(imgs=torch.rand(1, 3, 64, 64),
                                                                       masks=masks[ci, shots + li].reshape(1, 37, 37))})])
    m.postprocess_memory_negative()
    out = m(test_dict("test"))[0]
    rawn = ref_torch.RawBank(n_cls, l_neg, 1369, c)
    for li in range(l_neg):
        for ci in range(n_cls):
            ref_torch.bank_fill(rawn, [ci], feats[ci, shots + li][None], masks[ci, shots + li][None])
    _, neg_ins = ref_torch.bank_postprocess(rawn)
    ref = ref_torch.match_image(inp.lr_masks, inp.pred_ious, inp.tar_feat, pos_ins,
                                ref_torch.StageConfig(num_out_instance=10), (120, 160),
                                negative=dict(feats_avg=pos_avg, feats_ins_avg_neg=neg_ins))
    assert_same_ranking(out["scores"].cpu().numpy(), out["labels"].cpu().numpy(), ref["scores"].numpy(),
                        ref["labels"].numpy(), what="test with negatives")
    assert "memory_bank_neg.feats_ins_avg" in m.state_dict()
