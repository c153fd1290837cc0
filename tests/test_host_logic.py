"""CPU tests of the host-side logic: multi-rank memory-bank fill (gloo, world_size 2), image sharding, and the
model boundary's argument handling.  The pooling kernel itself needs a GPU, so these tests substitute the
oracle's arithmetic for `ops.fill_pool_accumulate` / `ops.fill_finalize` — the thing under test is the slot
assignment, the single all-reduce and the state-dict contract, not the kernel."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_pool(feat, soft_mask, enc_hw, sum_slot, wsum_slot, want_mask=False):
    m = torch.nn.functional.interpolate(soft_mask[None, None], size=tuple(enc_hw), mode="nearest").reshape(-1)
    sum_slot += (feat * m[:, None]).sum(0)
    wsum_slot += m.sum()
    return m


def _fake_finalize(sums, wsum):
    w = wsum.clone()
    w[w == 0] = 1.0
    wall = wsum.sum(1, keepdim=True)
    wall[wall == 0] = 1.0
    return sums / w[..., None], sums.sum(1) / wall


def _patch():
    ops = importlib.import_module("no-time-to-train_b200.ops")
    ops.fill_pool_accumulate = _fake_pool
    ops.fill_finalize = _fake_finalize


def _shots(n_cls, shots, c):
    synth = importlib.import_module("no-time-to-train_b200.synth")
    return synth.make_ref_shots(n_cls, shots, 1369, c, seed=5)


def _dataset_order(n_cls, shots):
    # COCOMemoryFillDataset emits exactly L consecutive items per category (coco_ref_dataset.py:348-361)
    return [(ci, li) for ci in range(n_cls) for li in range(shots)]


def _worker(rank, world, port, n_cls, shots, c, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _patch()
        pkg = importlib.import_module("no-time-to-train_b200")
        feats, masks = _shots(n_cls, shots, c)
        bank = pkg.MemoryBank(dict(category_num=n_cls, length=shots, feat_shape=(1369, c)))
        order = _dataset_order(n_cls, shots)
        # DistributedSampler(shuffle=False): rank r takes items r, r+W, ...
        for ci, li in order[rank::world]:
            bank.fill(ci, feats[ci, li], masks[ci, li].reshape(37, 37), (37, 37))
        sd = bank.state_dict()  # the pre-hook resolves the staged fill with one all_reduce
        bank.postprocess()
        torch.save(dict(sd={k: v.clone() for k, v in sd.items()}, ins=bank.feats_ins_avg, avg=bank.feats_avg,
                        counts=bank.fill_counts), os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_distributed_fill_matches_single_process_reference(tmp_path, world):
    from oracle import ref_torch
    n_cls, shots, c = 3, 2, 16
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, port, n_cls, shots, c, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    feats, masks = _shots(n_cls, shots, c)
    # reference semantics: every rank sees all gathered samples of a step in rank order (:471-485)
    raw = ref_torch.RawBank(n_cls, shots, 1369, c)
    order = _dataset_order(n_cls, shots)
    for step in range(len(order) // world):
        for r in range(world):
            ci, li = order[step * world + r]
            ref_torch.bank_fill(raw, [ci], feats[ci, li][None], masks[ci, li][None])
    want_avg, want_ins = ref_torch.bank_postprocess(raw)
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    for o in outs:
        assert torch.equal(o["counts"], raw.fill_counts)
        assert torch.equal(o["sd"]["masks"], raw.masks)
        np.testing.assert_allclose(o["ins"].numpy(), want_ins.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(o["avg"].numpy(), want_avg.numpy(), rtol=1e-5, atol=1e-6)
    # single-writer slots: the all-reduce is exact, so ranks agree bit for bit
    assert torch.equal(outs[0]["ins"], outs[1]["ins"]) and torch.equal(outs[0]["sd"]["feats_sum"], outs[1]["sd"]["feats_sum"])


def test_single_process_fill_and_state_dict_names():
    _patch()
    pkg = importlib.import_module("no-time-to-train_b200")
    n_cls, shots, c = 2, 2, 8
    feats, masks = _shots(n_cls, shots, c)
    bank = pkg.MemoryBank(dict(category_num=n_cls, length=shots, feat_shape=(1369, c)))
    for ci, li in [(0, 0), (1, 0), (0, 1)]:
        bank.fill(ci, feats[ci, li], masks[ci, li].reshape(37, 37), (37, 37))
    assert bank.fill_counts.tolist() == [2, 1]
    with pytest.raises(IndexError):
        bank.fill(0, feats[0, 0], masks[0, 0].reshape(37, 37), (37, 37))
    bank.postprocess()
    sd = bank.state_dict()
    for k in ("fill_counts", "masks", "feats_avg", "feats_ins_avg", "postprocessed"):
        assert k in sd  # the names the reference checkpoint uses (matching_baseline_utils.py:561-571)
    assert bool(sd["postprocessed"][0])
    # unfilled slot stays zero and takes part in the prototype mean
    assert float(bank.feats_ins_avg[1, 1].abs().max()) == 0.0
    # a reference-style checkpoint (extra raw `feats`, missing compact sums) loads with strict=False
    ref_sd = {k: v for k, v in sd.items() if k not in ("feats_sum", "mask_sum")}
    ref_sd["feats"] = torch.zeros(n_cls, shots, 4, c)
    other = pkg.MemoryBank(dict(category_num=n_cls, length=shots, feat_shape=(1369, c)))
    missing, unexpected = other.load_state_dict(ref_sd, strict=False)
    assert set(missing) == {"feats_sum", "mask_sum"} and unexpected == ["feats"]
    assert torch.equal(other.feats_ins_avg, bank.feats_ins_avg)


def test_image_sharding_is_strided_and_complete():
    """Test-time sharding mirrors DistributedSampler(shuffle=False) + the re-interleave of
    collect_results_cpu (run_lightning.py:69-75)."""
    shard = importlib.import_module("no-time-to-train_b200.sharding")
    for n, w in [(256, 8), (10, 4), (7, 2), (3, 4)]:
        parts = [shard.shard_indices(n, r, w) for r in range(w)]
        assert len({len(p) for p in parts}) == 1  # padded to equal length by repetition
        merged = shard.interleave([[f"img{i}" for i in p] for p in parts], n)
        assert merged == [f"img{i}" for i in range(n)]


def _collect_worker(rank, world, port, n_items, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = importlib.import_module("no-time-to-train_b200.sharding")
        # each rank "scores" its images: one encoded-result list per image, as `_output_inqueue` appends them
        part = [[dict(image_id=i, category_id=7, score=0.5, segmentation=dict(size=[4, 4], counts="`0"))]
                for i in shard.shard_indices(n_items, rank, world)]
        got = shard.collect_results(part, size=n_items)
        torch.save(got, os.path.join(out_dir, f"collect{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_collect_results_gathers_and_reorders(tmp_path):
    """`collect_results` == the reference's `collect_results_cpu` (run_lightning.py:23-78): rank 0 gets every image's
    results in dataset order, truncated to the dataset length (7 images on 2 ranks: one padded repeat); other
    ranks get None; without a process group the part comes back unchanged."""
    shard = importlib.import_module("no-time-to-train_b200.sharding")
    assert shard.collect_results([1, 2, 3], size=2) == [1, 2, 3]
    n_items, world = 7, 2
    port = 31500 + (os.getpid() % 2000)
    mp.start_processes(_collect_worker, args=(world, port, n_items, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    got0 = torch.load(os.path.join(tmp_path, "collect0.pt"))
    got1 = torch.load(os.path.join(tmp_path, "collect1.pt"))
    assert got1 is None
    assert [r[0]["image_id"] for r in got0] == list(range(n_items))


def test_model_rejects_unsupported_modes_without_gpu():
    pkg = importlib.import_module("no-time-to-train_b200")
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.Sam2MatchingBaselineNoAMG(sam2_infer_cfgs=dict(nms_thr=0.5, num_out_instance=10, cls_num_per_mask=1),
                                      memory_bank_cfg=dict(enable=True, category_num=2, length=1),
                                      encoder_geometry=(518, 14, 8))
