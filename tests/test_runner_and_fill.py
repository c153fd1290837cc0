"""The in-repo driver (`MatcherRunner`: the reference's setup / test_step / after_test, SURVEY.md §8b "who calls it"),
the batched memory-bank fill and the model's persistent output ring.

GPU tests run the three reference stages (fill_memory -> postprocess_memory -> test) through the drop-in class and
compare with the oracle; CPU tests (gloo, world size 2) cover the runner's sharding / gathering control flow with a
stand-in model."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_util import FILL_CASES, GOLDEN_DIR, assert_close_rel, assert_rows_match
from oracle import nttt_oracle as orc
from oracle import ref_torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


# ------------------------------------------------------------------------------------------------ CPU: control flow
class _FakeBank:
    def sync_fill(self):
        pass


class _FakeModel:
    """Stand-in for the drop-in class: records the modes it is called with, returns one instance per image."""
    training = False
    emit_rle = False

    def __init__(self):
        self.memory_bank = _FakeBank()
        self.calls = []

    def state_dict(self):
        return {"memory_bank.postprocessed": torch.ones(1, dtype=torch.bool)}

    def postprocess_memory(self):
        self.calls.append("postprocess")

    def __call__(self, batch):
        mode = batch[0]["data_mode"]
        self.calls.append(mode)
        if mode.startswith("fill"):
            return {}
        info = batch[0]["target_img_info"]
        m = torch.zeros(1, 4, 6, dtype=torch.bool)
        m[0, 1:3, 2:5] = True
        return [dict(binary_masks=m, bboxes=torch.tensor([[2, 1, 4, 2]]), scores=torch.tensor([0.5 + 0.001 * info["id"]]),
                     labels=torch.tensor([info["id"] % 3]), image_info=info)]


class _Images:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return dict(target_img=torch.zeros(3, 4, 6), target_img_info=dict(ori_height=4, ori_width=6, file_name=str(i), id=i))


def _runner_worker(rank, world, port, n_items, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = importlib.import_module("no-time-to-train_b200")
        model = _FakeModel()
        runner = pkg.MatcherRunner(model, "test", _Images(n_items), cat_inds_to_ids={0: 10, 1: 20, 2: 30}, rle=False)
        got = runner.run().after_test()
        torch.save(dict(got=got, calls=model.calls, n_times=len(runner.time_queue)), os.path.join(out_dir, f"run{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_runner_shards_and_gathers_like_the_reference(tmp_path):
    """7 images on 2 ranks: strided shards padded by repetition (DistributedSampler), results and per-image times
    re-interleaved on rank 0 and truncated to the dataset length (run_lightning.py:131-159)."""
    n_items, world = 7, 2
    port = 33500 + (os.getpid() % 2000)
    mp.start_processes(_runner_worker, args=(world, port, n_items, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    r0 = torch.load(os.path.join(tmp_path, "run0.pt"), weights_only=False)
    r1 = torch.load(os.path.join(tmp_path, "run1.pt"), weights_only=False)
    assert r1["got"] is None and r0["calls"] == ["test"] * 4 and r1["calls"] == ["test"] * 4
    res = r0["got"]
    assert [per_img[0]["image_id"] for per_img in res["results"]] == list(range(n_items))
    assert len(res["times"]) == n_items and len(res["results_unpacked"]) == n_items
    one = res["results_unpacked"][4]
    assert one["category_id"] == 20 and one["bbox"] == [2, 1, 2, 1]  # xywh without +1 (coco_ref_dataset.py:116-128)
    m = np.zeros((4, 6), bool)
    m[1:3, 2:5] = True
    assert one["segmentation"] == orc.encode_mask(m)  # host RLE mirror == the oracle's pycocotools restatement


def test_runner_modes_without_process_group():
    pkg = importlib.import_module("no-time-to-train_b200")
    model = _FakeModel()
    fills = [dict(refs_by_cat={0: dict(imgs=None, masks=None)}) for _ in range(3)]
    pkg.MatcherRunner(model, "fill_memory", fills).run().after_test()
    pkg.MatcherRunner(model, "postprocess_memory").run().after_test()
    assert model.calls == ["fill_memory"] * 3 + ["postprocess"]
    with pytest.raises(NotImplementedError, match="Unrecognized test mode"):
        pkg.MatcherRunner(model, "train")


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", FILL_CASES)
def test_fill_batch_equals_one_by_one_bit_for_bit(name):
    """A shot pools to the same bits alone (`fill`) or in a batch (`fill_batch`, one launch), whatever the batch
    split; both match the reference's golden fill."""
    P = importlib.import_module("no-time-to-train_b200")
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    n_cls, shots, c, seed, e_side, img_side = g["spec"].tolist()
    feats, _ = P.synth.make_ref_shots(n_cls, shots, e_side * e_side, c, seed=seed)
    order = g["order"].tolist()
    soft = torch.from_numpy(g["soft_masks"])
    cfg = dict(category_num=n_cls, length=shots, feat_shape=(e_side * e_side, c))
    solo = P.MemoryBank(cfg).to(DEV)
    for (ci, li), m in zip(order, soft):
        solo.fill(ci, feats[ci, li].to(DEV), m.to(DEV), (e_side, e_side))
    banks = [solo]
    for split in (len(order), 3):
        b = P.MemoryBank(cfg).to(DEV)
        for i in range(0, len(order), split):
            part = order[i:i + split]
            b.fill_batch([ci for ci, _ in part], torch.stack([feats[ci, li] for ci, li in part]).to(DEV),
                         soft[i:i + split].to(DEV), (e_side, e_side))
        banks.append(b)
    for b in banks:
        b.postprocess()
        assert np.array_equal(b.fill_counts.cpu().numpy(), g["fill_counts"])
        assert np.array_equal(b.masks.cpu().numpy(), g["masks_lowres"])
        assert_close_rel(b.feats_ins_avg.cpu().numpy(), g["feats_ins_avg"], what="feats_ins_avg")
        assert_close_rel(b.feats_avg.cpu().numpy(), g["feats_avg"], what="feats_avg")
    for b in banks[1:]:
        for k in ("feats_sum", "mask_sum", "masks", "feats_ins_avg", "feats_avg"):
            assert torch.equal(getattr(b, k), getattr(solo, k)), k
    with pytest.raises(IndexError):  # a batch that overflows a class is rejected before anything is written
        before = solo.feats_sum.clone()
        solo.fill_batch([0] * (shots + 1), feats[0, :1].repeat(shots + 1, 1, 1).to(DEV), soft[:1].repeat(shots + 1, 1, 1).to(DEV),
                        (e_side, e_side))
    assert torch.equal(solo.feats_sum, before)


@pytest.mark.gpu
def test_three_stages_through_the_runner_match_the_oracle(tmp_path):
    """fill_memory -> postprocess_memory -> test, each as its own runner over the reference's item dicts, with a
    checkpoint hand-off between the stages (run_lightning.py:107-119) and device-side RLE results."""
    P = importlib.import_module("no-time-to-train_b200")
    synth = P.synth
    n_cls, shots, c = 4, 3, 64
    feats, masks = synth.make_ref_shots(n_cls, shots, 1369, c, seed=29)
    images = [synth.make_stage_inputs(n=40, c=c, n_cls=n_cls, shots=shots, ori_hw=(240, 320), seed=300 + i,
                                      degenerate=2 if i == 1 else 0) for i in range(5)]

    class Model(P.Sam2MatchingBaselineNoAMG):
        def _forward_encoder(self, imgs):
            ci, li = self._next
            return feats[ci, li].to(imgs.device)[None]

        def _extract_target_features(self, tar_img, device):
            return images[self._cur].tar_feat.to(device), tar_img.to(device)

        def _forward_sam(self, imgs):
            inp = images[self._cur]
            return inp.lr_masks.to(imgs.device), inp.pred_ious.to(imgs.device), None

    def build():
        return Model(sam2_infer_cfgs=dict(points_per_side=32, testing_point_bs=256, iou_thr=0.4, nms_thr=0.5,
                                          num_out_instance=12, kmeans_k=2, n_pca_components=2, cls_num_per_mask=1),
                     memory_bank_cfg=dict(enable=True, category_num=n_cls, length=shots),
                     encoder_geometry=(518, 14, c), device=DEV)

    class Fill:
        order = [(ci, li) for ci in range(n_cls) for li in range(shots)]

        def __len__(self):
            return len(self.order)

        def __getitem__(self, i):
            ci, li = self.order[i]
            model._next = (ci, li)
            return dict(refs_by_cat={ci: dict(imgs=torch.rand(1, 3, 64, 64), masks=masks[ci, li].reshape(1, 37, 37))})

    class Test:
        def __len__(self):
            return len(images)

        def __getitem__(self, i):
            model._cur = i
            return dict(target_img=torch.zeros(3, 32, 32), target_img_info=dict(ori_height=240, ori_width=320,
                                                                                file_name=f"{i}.jpg", id=f"{i:06d}"))

    model = build()
    ck1, ck2 = str(tmp_path / "filled.ckpt"), str(tmp_path / "post.ckpt")
    P.MatcherRunner(model, "fill_memory", Fill()).run().after_test(ck1)
    model = build()  # a new process in the reference: the next stage starts from the checkpoint
    model.load_state_dict({k[len("seg_model."):]: v for k, v in torch.load(ck1)["state_dict"].items()}, strict=False)
    assert model.memory_bank.fill_counts.tolist() == [shots] * n_cls
    P.MatcherRunner(model, "postprocess_memory").run().after_test(ck2)
    model = build()
    model.load_state_dict({k[len("seg_model."):]: v for k, v in torch.load(ck2)["state_dict"].items()}, strict=False)
    runner = P.MatcherRunner(model, "test", Test(), cat_inds_to_ids={i: 100 + i for i in range(n_cls)})
    out = runner.run().after_test()
    assert len(out["results"]) == len(images) and len(out["times"]) == len(images)

    raw = ref_torch.RawBank(n_cls, shots, 1369, c)
    for ci, li in Fill.order:
        ref_torch.bank_fill(raw, [ci], feats[ci, li][None], masks[ci, li][None])
    _, want_ins = ref_torch.bank_postprocess(raw)
    assert_close_rel(model.memory_bank.feats_ins_avg.cpu().numpy(), want_ins.numpy(), what="feats_ins_avg")
    for i, per_img in enumerate(out["results"]):
        ref = ref_torch.match_image(images[i].lr_masks, images[i].pred_ious, images[i].tar_feat, want_ins,
                                    ref_torch.StageConfig(num_out_instance=12), (240, 320))
        assert len(per_img) == ref["scores"].shape[0] > 0
        got = dict(scores=np.array([r["score"] for r in per_img], np.float32),
                   labels=np.array([r["category_id"] - 100 for r in per_img]),
                   bboxes=np.array([[r["bbox"][0], r["bbox"][1], r["bbox"][0] + r["bbox"][2], r["bbox"][1] + r["bbox"][3]]
                                    for r in per_img]),
                   masks=np.stack([orc.rle_decode(orc.rle_from_string(r["segmentation"]["counts"].encode()), (240, 320))
                                   for r in per_img]))
        assert all(r["image_id"] == i and r["segmentation"]["size"] == [240, 320] for r in per_img)
        assert_rows_match(got, dict(scores=ref["scores"], labels=ref["labels"], bboxes=ref["bboxes"],
                                    masks=ref["binary_masks"]), what=f"runner image {i}")


@pytest.mark.gpu
def test_model_output_ring_equals_fresh_buffers():
    """forward_test writes `binary_masks` into a ring of persistent buffers; whatever was in a buffer before, a result
    equals the dense unpack into a fresh buffer, and stays intact for the next `output_ring - 1` calls."""
    P = importlib.import_module("no-time-to-train_b200")
    synth = P.synth
    n_cls, shots, c = 3, 2, 64
    images = [synth.make_stage_inputs(n=48, c=c, n_cls=n_cls, shots=shots, ori_hw=(200, 300), seed=500 + i,
                                      degenerate=2 if i % 2 else 1) for i in range(7)]

    class Model(P.Sam2MatchingBaselineNoAMG):
        def _extract_target_features(self, tar_img, device):
            return images[self._cur].tar_feat.to(device), tar_img.to(device)

        def _forward_sam(self, imgs):
            inp = images[self._cur]
            return inp.lr_masks.to(imgs.device), inp.pred_ious.to(imgs.device), None

    def build(ring):
        m = Model(sam2_infer_cfgs=dict(points_per_side=32, testing_point_bs=256, iou_thr=0.4, nms_thr=0.5,
                                       num_out_instance=9, kmeans_k=2, n_pca_components=2, cls_num_per_mask=1),
                  memory_bank_cfg=dict(enable=True, category_num=n_cls, length=shots),
                  encoder_geometry=(518, 14, c), device=DEV)
        m.output_ring = ring
        m.memory_bank.feats_ins_avg.copy_(images[0].feats_ins_avg)
        m.memory_bank.postprocessed[0] = True
        return m

    ringed, fresh = build(3), build(0)
    info = dict(ori_height=200, ori_width=300, file_name="x", id=0)
    held = []
    for i in range(len(images)):
        ringed._cur = fresh._cur = i
        a = ringed([dict(data_mode="test", target_img=torch.zeros(3, 8, 8), target_img_info=info)])[0]
        b = fresh([dict(data_mode="test", target_img=torch.zeros(3, 8, 8), target_img_info=info)])[0]
        for k in ("binary_masks", "bboxes", "labels"):
            assert torch.equal(a[k], b[k]), (i, k)
        assert torch.equal(torch.nan_to_num(a["scores"]), torch.nan_to_num(b["scores"]))
        held.append((a["binary_masks"], b["binary_masks"].clone()))
        for mine, want in held[-2:]:  # ring of 3: the previous result is still intact
            assert torch.equal(mine, want)
