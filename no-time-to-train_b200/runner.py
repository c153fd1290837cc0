"""Minimal driver of the matching model: the reference's three `run_lightning.py test` invocations without Lightning.

The reference drives `Sam2MatchingBaselineNoAMG` through `Sam2MatcherLightningModel`
(`no_time_to_train/pl_wrapper/sam2matcher_pl.py:163-239`: `setup`, `test_step`, bs=1 loader with the
`DistributedSampler` Lightning injects) and `SAM2RefLightningCLI` (`run_lightning.py:92-159`: `before_test`,
`after_test` = checkpoint save after the fill / post-process stages, cross-rank result collection and the FPS print
after the test stage).  pytorch_lightning, jsonargparse and mmengine are not part of this repository's environment, so
this module reproduces exactly those semantics for benchmarks and tests (SURVEY.md §8b "who calls it"):

    runner = MatcherRunner(model, "fill_memory", fill_dataset);     runner.run(); runner.after_test(out_path)
    runner = MatcherRunner(model, "postprocess_memory");            runner.run(); runner.after_test(out_path)
    runner = MatcherRunner(model, "test", test_dataset);            runner.run(); results = runner.after_test()

One process per GPU; with an initialised process group the items are sharded like the sampler does
(`sharding.shard_indices`) and results are gathered like `collect_results_cpu` does (`sharding.collect_results`).
The datasets yield the reference's own item dicts (`coco_ref_dataset.py:477-492`, `:758-807`): `refs_by_cat` for the
fill modes, `target_img` + `target_img_info` for the test modes; `data_mode` is set here if the item lacks it.
"""
from __future__ import annotations

import copy
import time

import torch
import torch.distributed as dist

from . import sharding
from .results import box_xyxy_to_xywh, rle_encode_host

FILL_MODES = ("fill_memory", "fill_memory_neg")
POST_MODES = ("postprocess_memory", "postprocess_memory_neg")
TEST_MODES = ("test", "test_support")


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def encode_output(output_dict, cat_inds_to_ids=None, segmentations=None):
    """`_output_inqueue` + `encode_results` (`sam2matcher_pl.py:144-158`, `coco_ref_dataset.py:590-613`) for one image:
    a list of COCO result dicts.  `segmentations` (COCO RLE dicts from the fused device-side encoder) replace the
    reference's dense-mask D2H copy + pycocotools call; without them the masks are encoded on the host by a plain
    column-major run-length pass (same wire format)."""
    info = output_dict["image_info"]
    img_id = int(info["id"]) if str(info["id"]).isdigit() else info["id"]
    scores = output_dict["scores"].cpu().tolist()
    labels = output_dict["labels"].cpu().tolist()
    boxes = output_dict["bboxes"].cpu().tolist()
    if segmentations is None:
        segmentations = [rle_encode_host(m) for m in output_dict["binary_masks"]]
    out = []
    for score, label, box, seg in zip(scores, labels, boxes, segmentations):
        cat = int(label) if cat_inds_to_ids is None else int(cat_inds_to_ids[int(label)])
        out.append(dict(image_id=img_id, category_id=cat, bbox=box_xyxy_to_xywh(box), score=float(score),
                        segmentation=seg))
    return out


class MatcherRunner:
    """`Sam2MatcherLightningModel.setup/test_step` + `SAM2RefLightningCLI.after_test` for one test_mode."""

    def __init__(self, seg_model, test_mode: str, dataset=None, cat_inds_to_ids=None, rle: bool = True):
        if test_mode not in FILL_MODES + POST_MODES + TEST_MODES:
            raise NotImplementedError("Unrecognized test mode: %s" % test_mode)  # sam2matcher_pl.py:198
        self.seg_model = seg_model
        self.test_mode = test_mode
        self.eval_dataset = dataset if dataset is not None else [None]  # DummyDataset(1) for the post-process modes
        self.cat_inds_to_ids = cat_inds_to_ids
        self.rle = rle
        if rle and hasattr(seg_model, "emit_rle"):
            seg_model.emit_rle = True  # the model adds `segmentations` (device-side COCO RLE) to its output dict
        self.setup()

    def setup(self):
        """(`sam2matcher_pl.py:203-214`) result queues."""
        self.output_queue = []
        self.time_queue = []

    # -------------------------------------------------------------------------------------------- per item
    def test_step(self, batch):
        """(`sam2matcher_pl.py:163-201`) `batch` is the bs=1 identity-collated list of one item dict."""
        assert not self.seg_model.training
        with torch.inference_mode():
            mode = self.test_mode
            if mode in FILL_MODES:
                batch[0].setdefault("data_mode", mode)
                self.seg_model(batch)
            elif mode == "postprocess_memory":
                self.seg_model.postprocess_memory()
            elif mode == "postprocess_memory_neg":
                self.seg_model.postprocess_memory_negative()
            elif mode == "test_support":
                batch[0].setdefault("data_mode", mode)
                output = self.seg_model(batch)
                assert len(output) == len(batch)
                self._output_inqueue(output[0])
            else:  # "test": the reference brackets the forward with device synchronisations and wall-clock time
                batch[0].setdefault("data_mode", mode)
                if torch.cuda.is_available():  # (as :178-181)
                    torch.cuda.synchronize()
                start_time = time.time()
                output = self.seg_model(batch)
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                self.time_queue.append(time.time() - start_time)
                assert len(output) == len(batch)
                self._output_inqueue(output[0])

    def _output_inqueue(self, output_dict):
        segs = output_dict.get("segmentations") if self.rle else None
        self.output_queue.append(encode_output(output_dict, self.cat_inds_to_ids, segs))

    # -------------------------------------------------------------------------------------------- whole stage
    def run(self):
        """`trainer.test`: this rank's share of the dataset, in the sampler's order."""
        rank, world = _rank_world()
        for idx in sharding.shard_indices(len(self.eval_dataset), rank, world):
            item = self.eval_dataset[idx]
            self.test_step([dict(item) if item is not None else None])
        return self

    def after_test(self, out_path=None):
        """(`run_lightning.py:99-186`).  Fill / post-process modes: resolve the distributed fill (collective) and save
        the checkpoint — every rank builds the state dict, rank 0 writes it, as `trainer.save_checkpoint` does.
        Test modes: gather results and per-image times on rank 0 (`collect_results_cpu`), print the reference's
        FPS lines and return `dict(results=..., results_unpacked=..., times=...)` (None on the other ranks)."""
        rank, world = _rank_world()
        if self.test_mode in FILL_MODES + POST_MODES:
            bank = self.seg_model.memory_bank if not self.test_mode.endswith("_neg") else self.seg_model.memory_bank_neg
            bank.sync_fill()
            state = {"state_dict": {"seg_model." + k: v for k, v in self.seg_model.state_dict().items()}}
            if out_path is not None and rank == 0:
                torch.save(state, out_path)
            return state
        n = len(self.eval_dataset)
        results_all = sharding.collect_results(copy.deepcopy(self.output_queue), size=n)
        times_all = sharding.collect_results(copy.deepcopy(self.time_queue), size=n) if self.time_queue else None
        if rank != 0:
            return None
        if times_all:
            total = float(sum(times_all))
            print("\n[Validation] Inference Time Benchmark:")
            print(f"  Total images: {len(times_all)}")
            print(f"  Total time: {total:.4f} s")
            print(f"  Average time per image: {total / len(times_all):.4f} s")
            print(f"  FPS: {len(times_all) / total:.2f}")
        unpacked = []
        for per_img in results_all:
            unpacked.extend(per_img)
        return dict(results=results_all, results_unpacked=unpacked, times=times_all)
