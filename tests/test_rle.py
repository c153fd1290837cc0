"""COCO run-length encoding of the result masks.

CPU: the C restatement of pycocotools' rleEncode against counts produced by the reference's own
`mask_to_rle_pytorch` (sam2/utils/amg.py:111-140; tests/golden/rle_counts.npz), and the rleToString restatement
through its structural properties (pycocotools itself is not in this image: the string half is unpinned).
GPU: `nttt_rle_encode` / the fused stage output against the oracle, bit-exact counts and strings."""
import importlib
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, assert_rows_match, load_case
from oracle import nttt_oracle as orc

DEV = "cuda:0"
G = dict(np.load(os.path.join(GOLDEN_DIR, "rle_counts.npz")))
SPECIAL = sorted({k.split("__")[1] for k in G if k.startswith("special__")})
STAGE = sorted({k.split("__")[1] for k in G if k.startswith("stage__")})


def _special(name):
    h, w = G[f"special__{name}__hw"].tolist()
    mask = np.unpackbits(G[f"special__{name}__mask"])[:h * w].reshape(h, w).astype(bool)
    return mask, G[f"special__{name}__counts"]


def _stage_masks(case):
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    oh, ow = int(g["spec"][4]), int(g["spec"][5])
    k = g["out_masks_packed"].shape[0]
    masks = np.unpackbits(g["out_masks_packed"], axis=-1)[:, :oh * ow].reshape(k, oh, ow).astype(bool)
    lens = G[f"stage__{case}__lens"]
    flat = G[f"stage__{case}__counts"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    return masks, [flat[offs[i]:offs[i + 1]] for i in range(k)]


# ------------------------------------------------------------------------------------------------ CPU (oracle)
@pytest.mark.parametrize("name", SPECIAL)
def test_oracle_counts_match_reference_special(name):
    mask, want = _special(name)
    got = orc.rle_counts(mask)
    assert np.array_equal(got.astype(np.int64), want)
    assert np.array_equal(orc.rle_decode(got, mask.shape), mask)


@pytest.mark.parametrize("case", STAGE)
def test_oracle_counts_match_reference_stage_outputs(case):
    masks, wants = _stage_masks(case)
    for m, want in zip(masks, wants):
        assert np.array_equal(orc.rle_counts(m).astype(np.int64), want)


def test_oracle_string_roundtrip_and_alphabet():
    rng = np.random.default_rng(3)
    for hw in [(1, 1), (7, 5), (64, 64), (333, 500), (1024, 1024)]:
        dens = 0.5 if hw[0] * hw[1] < 10000 else 0.02
        m = rng.random(hw) < dens
        c = orc.rle_counts(m)
        s = orc.rle_to_string(c)
        assert set(s) <= set(range(48, 112))  # "ascii chars 48-111"
        assert np.array_equal(orc.rle_from_string(s), c)
    # hand-checked encodings of single counts: 16 -> 0b10000 needs a second group because bit 4 is the sign bit
    assert orc.rle_to_string(np.array([16], np.uint32)) == b"`0"
    assert orc.rle_to_string(np.array([0, 16], np.uint32)) == b"0`0"
    assert orc.rle_to_string(np.array([5], np.uint32)) == b"5"
    # the 4th count is stored as a difference to the 2nd: 3 - 7 = -4 -> 0b11100 -> one char, 28 + 48
    assert orc.rle_to_string(np.array([1, 7, 2, 3], np.uint32)) == b"172" + bytes([28 + 48])


def _kernel_walk(mask, rect):
    """Host restatement of the WALK `rle_encode_kernel` (csrc/rle.cu) makes over a mask: only rows [r0, min(r1+1, oh))
    and columns [32*w0, min(32*w1+1, ow)) are visited, nothing outside the rect is loaded, the carry into the first row
    of a column is the wrap from the previous column only when the rect starts at row 0, and a rect that reaches the last
    row without reaching the first owes the boundary at (x, 0) explicitly (`rle_wrap_edge`)."""
    oh, ow = mask.shape
    r0, r1, w0, w1 = rect
    if r1 <= r0 or w1 <= w0:
        return np.array([oh * ow], dtype=np.int64)
    rend, cend = min(r1 + 1, oh), min((w1 << 5) + 1, ow)

    def px(y, x):
        inside = r0 <= y < r1 and w0 <= (x >> 5) < w1 and x < ow
        return int(mask[y, x]) if inside else 0

    def prev_column_last(x):
        if x == 0 or r1 < oh or not w0 <= ((x - 1) >> 5) < w1:
            return 0
        return int(mask[oh - 1, x - 1])

    edges = []
    for x in range(w0 << 5, cend):
        if r0 > 0 and prev_column_last(x):
            edges.append(x * oh)
        carry = prev_column_last(x) if r0 == 0 else 0
        for y in range(r0, rend):
            v = px(y, x)
            if v != carry:
                edges.append(x * oh + y)
            carry = v
    return np.diff(np.concatenate([[0], np.array(edges + [oh * ow], dtype=np.int64)]))


@pytest.mark.parametrize("name", SPECIAL)
def test_kernel_walk_over_tight_and_full_rects(name):
    """The rect-confined walk the CUDA kernel makes gives the reference's counts for tight and full rects (pins the
    bottom-border case: a tight rect with r1 == oh and r0 > 0)."""
    mask, want = _special(name)
    h, w = mask.shape
    for rect in (_tight_rect(mask), [0, h, 0, (w + 31) // 32]):
        assert np.array_equal(_kernel_walk(mask, rect), want), (name, rect)


# ------------------------------------------------------------------------------------------------ GPU
def _pack_rows(mask):
    """[h, w] bool -> packed words [h, words] int32 (bit b of word w = pixel 32w+b), as the stage stores them."""
    h, w = mask.shape
    words = (w + 31) // 32
    padded = np.zeros((h, words * 32), dtype=np.uint8)
    padded[:, :w] = mask
    return np.packbits(padded.reshape(h, words, 32), axis=-1, bitorder="little").view(np.uint32).reshape(h, words)


def _tight_rect(mask):
    ys, xs = np.nonzero(mask)
    if ys.size == 0:
        return [0, 0, 0, 0]
    return [int(ys.min()), int(ys.max()) + 1, int(xs.min()) // 32, int(xs.max()) // 32 + 1]


def _encode_on_gpu(ops, masks, rects, cap_counts=16384, cap_chars=32768, slot=None, count=None):
    k = len(masks)
    h, w = masks[0].shape
    bits = torch.from_numpy(np.stack([_pack_rows(m) for m in masks]).view(np.int32)).to(DEV)
    rect = torch.tensor(rects, dtype=torch.int32, device=DEV)
    cnt = torch.tensor([k if count is None else count], dtype=torch.int32, device=DEV)
    slot_t = None if slot is None else torch.tensor(slot, dtype=torch.int32, device=DEV)
    counts, n_counts, chars, n_chars = ops.rle_encode(bits, rect, slot_t, cnt, (h, w), cap_counts, cap_chars,
                                                      max_count=k if slot is None else len(slot))
    return (counts.cpu().numpy().view(np.uint32), n_counts.cpu().numpy(), chars.cpu().numpy(), n_chars.cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("name", SPECIAL)
@pytest.mark.parametrize("rect_kind", ["tight", "full"])
def test_gpu_rle_special_masks(name, rect_kind):
    ops = importlib.import_module("no-time-to-train_b200.ops")
    mask, want = _special(name)
    h, w = mask.shape
    rect = _tight_rect(mask) if rect_kind == "tight" else [0, h, 0, (w + 31) // 32]
    counts, n_counts, chars, n_chars = _encode_on_gpu(ops, [mask], [rect])
    m = int(n_counts[0])
    assert np.array_equal(counts[0, :m].astype(np.int64), want)
    assert chars[0, :n_chars[0]].tobytes() == orc.rle_to_string(want.astype(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("case", STAGE)
def test_gpu_rle_stage_outputs(case):
    ops = importlib.import_module("no-time-to-train_b200.ops")
    masks, wants = _stage_masks(case)
    order = list(range(len(masks)))[::-1]  # encode through a slot indirection, reversed
    counts, n_counts, chars, n_chars = _encode_on_gpu(ops, list(masks), [_tight_rect(m) for m in masks], slot=order)
    for j, k in enumerate(order):
        m = int(n_counts[j])
        assert np.array_equal(counts[j, :m].astype(np.int64), wants[k]), (case, k)
        assert chars[j, :n_chars[j]].tobytes() == orc.rle_to_string(wants[k].astype(np.uint32))


@pytest.mark.gpu
def test_gpu_rle_large_random_and_overflow():
    ops = importlib.import_module("no-time-to-train_b200.ops")
    rng = np.random.default_rng(11)
    big = np.zeros((1024, 1024), dtype=bool)
    big[100:900, 130:777] = rng.random((800, 647)) < 0.97  # speckled blob: tens of thousands of runs
    full = np.ones((1024, 1024), dtype=bool)
    masks = [big, full, np.zeros((1024, 1024), dtype=bool)]
    rects = [_tight_rect(big), [0, 1024, 0, 32], [0, 0, 0, 0]]
    want = [orc.rle_counts(m) for m in masks]
    cap = int(len(want[0])) + 8
    counts, n_counts, chars, n_chars = _encode_on_gpu(ops, masks, rects, cap_counts=cap, cap_chars=4 * cap)
    for j in range(3):
        assert int(n_counts[j]) == len(want[j])
        assert np.array_equal(counts[j, :n_counts[j]], want[j])
        assert chars[j, :n_chars[j]].tobytes() == orc.rle_to_string(want[j])
    # live count below the capacity: dead rows report zero
    counts, n_counts, chars, n_chars = _encode_on_gpu(ops, masks, rects, cap_counts=cap, cap_chars=4 * cap, count=1)
    assert n_counts.tolist()[1:] == [0, 0] and n_chars.tolist()[1:] == [0, 0]
    # overflow of the counts capacity: size needed is reported, -1 chars
    counts, n_counts, chars, n_chars = _encode_on_gpu(ops, masks, rects, cap_counts=1000, cap_chars=4000)
    assert int(n_counts[0]) == len(want[0]) and int(n_chars[0]) == -1
    assert int(n_counts[1]) == 2 and chars[1, :n_chars[1]].tobytes() == orc.rle_to_string(want[1])
    # overflow of the string capacity only: true length reported
    counts, n_counts, chars, n_chars = _encode_on_gpu(ops, masks, rects, cap_counts=cap, cap_chars=100)
    assert int(n_chars[0]) == len(orc.rle_to_string(want[0])) > 100


@pytest.mark.gpu
def test_gpu_rle_compact_packs_strings_back_to_back():
    """nttt_rle_compact: string j at offset sum_{i<j} len_i, overflowed rows (len < 0 or > cap) skipped, nothing
    written past the output capacity."""
    ops = importlib.import_module("no-time-to-train_b200").ops
    g = torch.Generator().manual_seed(5)
    k, cap = 37, 512
    chars = torch.randint(48, 112, (k, cap), dtype=torch.uint8, generator=g)
    lens = torch.randint(0, cap + 1, (k,), dtype=torch.int32, generator=g)
    lens[3], lens[11], lens[20], lens[21] = -1, cap + 9, 0, cap
    want = b"".join(chars[j, :int(lens[j])].numpy().tobytes() for j in range(k) if 0 <= int(lens[j]) <= cap)
    out = ops.rle_compact(chars.cuda(), lens.cuda(), k)
    assert out[:len(want)].cpu().numpy().tobytes() == want
    # a prefix only, into an output with room for the first five strings: the sixth must not be written
    room = sum(int(lens[j]) for j in range(5) if 0 <= int(lens[j]) <= cap)
    small = torch.full((room + 64,), 7, dtype=torch.uint8, device="cuda")
    ops.rle_compact(chars.cuda(), lens.cuda(), 5, out=small[:room])
    assert small[:room].cpu().numpy().tobytes() == want[:room] and bool((small[room:] == 7).all())
    tiny = torch.full((8,), 7, dtype=torch.uint8, device="cuda")
    ops.rle_compact(chars.cuda(), lens.cuda(), k, out=tiny[:4])
    assert bool((tiny[4:] == 7).all())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["stage_a_1024_degenerate", "stage_f_truncate_333x500", "stage_d_200x180_downscale",
                                  "stage_h_borders_640x480", "stage_i_borders_1024"])
def test_stage_rle_output_matches_dense_masks(name):
    """Fused: the stage's RLE strings decode to exactly its own dense masks and equal the oracle's encoding of the
    reference's masks; with dense_masks=False the result is the same without producing the bool masks."""
    P = importlib.import_module("no-time-to-train_b200")
    g, inp, cfg = load_case(name)
    stage = P.MatchingStage(DEV, P.StageConfig(nms_thr=0.5, num_out_instance=cfg["num_out_instance"], enc_hw=(37, 37)))
    stage.set_prototypes(inp.feats_ins_avg)
    args = (inp.lr_masks.to(DEV), inp.pred_ious.to(DEV), inp.tar_feat.to(DEV), inp.ori_hw)
    out = stage.match(*args, rle=True)
    segs = out["segmentations"]
    dense = out["binary_masks"].cpu().numpy()
    assert len(segs) == dense.shape[0] > 0
    for j, seg in enumerate(segs):
        assert seg["size"] == list(inp.ori_hw)
        assert seg == orc.encode_mask(dense[j])
        assert np.array_equal(orc.rle_decode(orc.rle_from_string(seg["counts"].encode()), inp.ori_hw), dense[j])
    lean = stage.match(*args, rle=True, dense_masks=False)
    assert lean["binary_masks"] is None and lean["segmentations"] == segs
    assert torch.equal(lean["bboxes"], out["bboxes"]) and torch.equal(lean["labels"], out["labels"])
    # against the reference's masks (golden): every output's string equals the oracle's encoding of the reference mask
    # it matches (rows may be permuted inside float score ties only)
    oh, ow = inp.ori_hw
    ref_masks = np.unpackbits(g["out_masks_packed"], axis=-1)[:, :oh * ow].reshape(-1, oh, ow).astype(bool)
    assert_rows_match(dict(scores=out["scores"], labels=out["labels"], bboxes=out["bboxes"], masks=dense),
                      dict(scores=g["out_scores"], labels=g["out_labels"], bboxes=g["out_bboxes"], masks=ref_masks),
                      what=name)
    ref_segs = [orc.encode_mask(m) for m in ref_masks]
    assert sorted(s["counts"] for s in segs) == sorted(s["counts"] for s in ref_segs)


@pytest.mark.gpu
def test_encode_results_mirrors_reference_dicts():
    """`encode_results` (coco_ref_dataset.py:590-613) from the fused device-side RLE: same keys, xywh boxes, category
    mapping, and segmentations that equal the oracle's encoding of the dense masks."""
    P = importlib.import_module("no-time-to-train_b200")
    g, inp, cfg = load_case("stage_b_480x640")
    stage = P.MatchingStage(DEV, P.StageConfig(nms_thr=0.5, num_out_instance=cfg["num_out_instance"], enc_hw=(37, 37)))
    stage.set_prototypes(inp.feats_ins_avg)
    pend = stage.match_async(inp.lr_masks.to(DEV), inp.pred_ious.to(DEV), inp.tar_feat.to(DEV), inp.ori_hw, rle=True)
    cat_map = {i: 100 + 3 * i for i in range(cfg["n_cls"])}
    res = P.encode_results(pend, "000123", cat_map)
    out = pend.get()
    dense = out["binary_masks"].cpu().numpy()
    assert len(res) == dense.shape[0]
    for j, r in enumerate(res):
        assert set(r) == {"image_id", "category_id", "bbox", "score", "segmentation"}
        assert r["image_id"] == 123
        assert r["category_id"] == cat_map[int(out["labels"][j])]
        x1, y1, x2, y2 = out["bboxes"][j].tolist()
        assert r["bbox"] == [x1, y1, x2 - x1, y2 - y1]
        assert r["segmentation"] == orc.encode_mask(dense[j])
