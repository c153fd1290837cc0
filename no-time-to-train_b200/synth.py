"""Seeded synthetic inputs at the `_forward_sam` / `_extract_target_features` seam.

The reference's matching stage starts after the frozen encoders
(`no_time_to_train/models/Sam2MatchingBaseline_noAMG.py:576-580`): it receives `tar_feat [E,C]`,
`lr_masks [N,256,256]` (mask logits) and `pred_ious [N]`.  Random-init SAM-2 emits degenerate masks, so
tests, golden vectors and the benchmark inject synthetic tensors at that seam.

Everything is generated with a CPU `torch.Generator`, so the same seed yields the same tensors in the
authoring container and on the GPU box (same torch build in both).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

LOWRES = 256  # SAM-2 low-res mask side (sam2/modeling/sam/mask_decoder.py:189-286 emits 256x256 logits)


@dataclass
class StageInputs:
    lr_masks: torch.Tensor       # [N,256,256] f32 logits
    pred_ious: torch.Tensor      # [N] f32
    tar_feat: torch.Tensor       # [E,C] f32
    feats_ins_avg: torch.Tensor  # [n_cls,L,C] f32 (memory bank, post-processed)
    ori_hw: tuple                # (ori_height, ori_width)


def make_masks(n: int, gen: torch.Generator, lo: int = LOWRES, noise: float = 0.5,
               rmin: float = 4.0, rmax: float = 64.0) -> torch.Tensor:
    """N elliptical blobs: logit = 8(1 - (dy/ry)^2 - (dx/rx)^2) + noise*N(0,1)."""
    cy = torch.rand(n, generator=gen) * lo
    cx = torch.rand(n, generator=gen) * lo
    lr = math.log(rmin)
    hr = math.log(rmax)
    ry = torch.exp(torch.rand(n, generator=gen) * (hr - lr) + lr)
    rx = torch.exp(torch.rand(n, generator=gen) * (hr - lr) + lr)
    ys = torch.arange(lo, dtype=torch.float32).view(1, lo, 1)
    xs = torch.arange(lo, dtype=torch.float32).view(1, 1, lo)
    dy = (ys - cy.view(n, 1, 1)) / ry.view(n, 1, 1)
    dx = (xs - cx.view(n, 1, 1)) / rx.view(n, 1, 1)
    logits = 8.0 * (1.0 - dy * dy - dx * dx)
    logits += noise * torch.randn(n, lo, lo, generator=gen)
    return logits.contiguous()


def make_features(e: int, c: int, n_cls: int, shots: int, gen: torch.Generator,
                  clustered: bool = True, n_centres: int = 5):
    """Target features [E,C] and a post-processed bank feats_ins_avg [n_cls,L,C].

    clustered=True follows SURVEY.md §8d's second scenario: prototypes are `n_centres` random unit
    vectors plus small noise tiled over the classes and target features are drawn near them, so labels
    concentrate on a few classes, class-aware NMS actually fires and IoS groups are large.
    """
    if clustered:
        centres = torch.nn.functional.normalize(torch.randn(n_centres, c, generator=gen), dim=-1)
        cls_centre = torch.arange(n_cls) % n_centres
        bank = centres[cls_centre].unsqueeze(1) + (0.3 / math.sqrt(c)) * torch.randn(n_cls, shots, c, generator=gen)
        patch_centre = torch.randint(0, n_centres, (e,), generator=gen)
        tar = centres[patch_centre] + (0.6 / math.sqrt(c)) * torch.randn(e, c, generator=gen)
    else:
        bank = torch.randn(n_cls, shots, c, generator=gen)
        tar = torch.randn(e, c, generator=gen)
    return tar.contiguous(), bank.contiguous()


def make_stage_inputs(n: int, c: int, n_cls: int, shots: int, ori_hw=(1024, 1024), seed: int = 1234,
                      e_side: int = 37, clustered: bool = True, degenerate=False) -> StageInputs:
    """degenerate: False/0 = plain blobs; True/1 = + the edge cases of `inject_degenerate_cases`;
    2 = + masks cut by exactly one image border (`inject_border_cases`)."""
    gen = torch.Generator().manual_seed(seed)
    lr_masks = make_masks(n, gen)
    pred_ious = 0.4 + 0.6 * torch.rand(n, generator=gen)
    tar_feat, bank = make_features(e_side * e_side, c, n_cls, shots, gen, clustered=clustered)
    if degenerate:
        inject_degenerate_cases(lr_masks, bank)
    if int(degenerate) >= 2:
        inject_border_cases(lr_masks, pred_ious)
    return StageInputs(lr_masks, pred_ious, tar_feat, bank, tuple(ori_hw))


def inject_degenerate_cases(lr_masks: torch.Tensor, bank: torch.Tensor) -> None:
    """Overwrite a few masks / bank slots with the edge cases SURVEY.md §8c lists (needs N >= 8)."""
    n = lr_masks.shape[0]
    assert n >= 8
    # 0: empty low-res mask (all logits negative) -> zero feature row, all sims 0, score filtered out
    lr_masks[0] = -3.0
    # 1: one-pixel-wide vertical line -> zero-area box (x1 == x2) in NMS
    lr_masks[1] = -5.0
    lr_masks[1, 40:90, 77] = 4.0
    # 2: a single weakly positive pixel in a very negative field: non-empty at 256^2, but the
    #    antialiased bilinear resize pulls every full-res sample below zero -> empty full-res mask
    lr_masks[2] = -50.0
    lr_masks[2, 128, 128] = 0.5
    # 3 and 4: identical masks (tied boxes; suppression must keep the higher pred_iou / lower index)
    lr_masks[4] = lr_masks[3]
    # 5: mask touching the image border on all sides
    lr_masks[5] = 2.0
    # 6: exact zeros are NOT foreground (strict > 0)
    lr_masks[6] = 0.0
    lr_masks[6, 10:20, 10:20] = 1.0
    # an unfilled prototype slot: reference averages over ALL L slots, zeros included
    if bank.shape[1] > 1:
        bank[0, -1] = 0.0


def inject_border_cases(lr_masks: torch.Tensor, pred_ious: torch.Tensor) -> None:
    """Masks cut by exactly ONE image border (needs N >= 16): their full-resolution rect reaches the last row / column
    without reaching the first one, the case in which a column-major run wraps into a column whose top is background.
    Their SAM scores are raised so that they survive NMS and reach the output."""
    n = lr_masks.shape[0]
    assert n >= 16
    def block(i, ys, xs):
        lr_masks[i] = -4.0
        lr_masks[i, ys[0]:ys[1], xs[0]:xs[1]] = 3.0
        pred_ious[i] = 0.99 - 0.001 * i
    block(8, (200, 256), (60, 100))     # bottom border only, several columns
    block(9, (180, 256), (130, 131))    # bottom border only, one low-res column
    block(10, (230, 256), (220, 256))   # bottom-right corner
    block(11, (0, 30), (100, 140))      # top border only
    block(12, (90, 140), (236, 256))    # right border only
    block(13, (150, 256), (0, 24))      # bottom-left corner
    # a bottom-border mask with a ragged lower edge: columns alternate between reaching the last row and not
    lr_masks[14] = -4.0
    lr_masks[14, 210:250, 150:200] = 3.0
    lr_masks[14, 250:256, 150:200:2] = 3.0
    pred_ious[14] = 0.97


def make_ref_shots(n_cls: int, shots: int, e: int, c: int, seed: int = 4321):
    """Synthetic reference shots at the encoder-output seam of `forward_fill_memory`
    (`Sam2MatchingBaseline_noAMG.py:461-469`): per (class, shot) patch features [E,C] and a SOFT
    mask [E] in [0,1] (reference masks are bilinear-resized then nearest-sampled)."""
    gen = torch.Generator().manual_seed(seed)
    feats = torch.randn(n_cls, shots, e, c, generator=gen)
    side = int(round(math.sqrt(e)))
    masks = torch.zeros(n_cls, shots, side, side)
    for ci in range(n_cls):
        for li in range(shots):
            y0, x0 = torch.randint(0, side // 2, (2,), generator=gen).tolist()
            hh, ww = torch.randint(3, side // 2, (2,), generator=gen).tolist()
            masks[ci, li, y0:y0 + hh, x0:x0 + ww] = 1.0
            # soft rim
            masks[ci, li, y0, x0:x0 + ww] = 0.5
            masks[ci, li, y0:y0 + hh, x0] = 0.25
    return feats, masks.reshape(n_cls, shots, e)


def make_multimask_inputs(n: int, m: int = 4, seed: int = 77, jitter: float = 0.35):
    """Synthetic RAW decoder output at the `sam_mask_decoder` seam (`Sam2MatchingBaseline_noAMG.py:275-288`):
    `multi [n, m, 256, 256]` logits (m jittered variants of one blob per prompt) and `ious [n, m]` in (0.2, 1).
    A few rows carry the edge cases of the selection: a tie between two planes (first wins), plane 0 holding the
    largest IoU (it never competes), an IoU exactly at a typical threshold."""
    gen = torch.Generator().manual_seed(seed)
    base = make_masks(n, gen)
    multi = torch.empty((n, m, LOWRES, LOWRES), dtype=torch.float32)
    for j in range(m):
        shift = jitter * (j + 1) * torch.randn(n, 1, 1, generator=gen)
        multi[:, j] = base * (1.0 + 0.1 * j) + shift + 0.3 * torch.randn(n, LOWRES, LOWRES, generator=gen)
    ious = 0.2 + 0.8 * torch.rand(n, m, generator=gen)
    if n >= 6 and m >= 3:
        ious[0, 1] = ious[0, 2] = 0.9   # tie: plane 1 wins
        ious[0, 3:] = 0.3
        ious[1, 0] = 0.99               # plane 0 is skipped even when it is the best
        ious[2, 1:] = 0.5               # all equal to the threshold used by the tests: filtered (strict >)
    return multi.contiguous(), ious.contiguous()


# ----------------------------------------------------------------------------------------------------------
# Device-side generators (bench.py): the same recipes with a CUDA generator, so that hundreds of images / thousands
# of reference shots can be produced in milliseconds each.  Seeds are per item (image index, (class, shot)), so every
# rank can regenerate any item; values differ from the CPU generator's (different RNG stream), the distributions do not.
# ----------------------------------------------------------------------------------------------------------
def cluster_centres(c: int, n_centres: int = 5, seed: int = 99) -> torch.Tensor:
    """The `n_centres` unit vectors shared by reference shots and target features of one synthetic dataset."""
    gen = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n_centres, c, generator=gen), dim=-1)


def make_stage_inputs_device(n: int, centres: torch.Tensor, device, seed: int, e_side: int = 37, noise: float = 0.5,
                             rmin: float = 4.0, rmax: float = 64.0):
    """One synthetic image at the encoder seams, generated on `device`: (lr_masks [n,256,256], pred_ious [n],
    tar_feat [E, C]) following `make_masks` / `make_features(clustered=True)`."""
    gen = torch.Generator(device=device).manual_seed(seed)
    lo = LOWRES
    r = torch.rand((5, n), generator=gen, device=device)
    cy, cx = r[0] * lo, r[1] * lo
    lr_, hr_ = math.log(rmin), math.log(rmax)
    ry, rx = torch.exp(r[2] * (hr_ - lr_) + lr_), torch.exp(r[3] * (hr_ - lr_) + lr_)
    ys = torch.arange(lo, dtype=torch.float32, device=device).view(1, lo, 1)
    xs = torch.arange(lo, dtype=torch.float32, device=device).view(1, 1, lo)
    dy = (ys - cy.view(n, 1, 1)) / ry.view(n, 1, 1)
    dx = (xs - cx.view(n, 1, 1)) / rx.view(n, 1, 1)
    logits = 8.0 * (1.0 - dy * dy - dx * dx)
    logits += noise * torch.randn((n, lo, lo), generator=gen, device=device)
    pred_ious = 0.4 + 0.6 * r[4]
    e, c = e_side * e_side, centres.shape[1]
    cen = centres.to(device)
    patch_centre = torch.randint(0, cen.shape[0], (e,), generator=gen, device=device)
    tar = cen[patch_centre] + (0.6 / math.sqrt(c)) * torch.randn((e, c), generator=gen, device=device)
    return logits.contiguous(), pred_ious.contiguous(), tar.contiguous()


def make_ref_shot_device(cls: int, shot: int, centres: torch.Tensor, device, e_side: int = 37, mask_side: int = 518,
                         seed: int = 4321):
    """One synthetic reference shot at the encoder-output seam of `forward_fill_memory`
    (`Sam2MatchingBaseline_noAMG.py:461-469`): patch features [E, C] near the class's cluster centre and a SOFT mask
    [mask_side, mask_side] in [0,1] (a box with a soft rim), reproducible from (cls, shot) alone."""
    gen = torch.Generator(device=device).manual_seed(seed + 100003 * cls + shot)
    e, c = e_side * e_side, centres.shape[1]
    cen = centres.to(device)[cls % centres.shape[0]]
    feats = cen.unsqueeze(0) + (3.0 / math.sqrt(c)) * torch.randn((e, c), generator=gen, device=device)
    geo = torch.rand((4,), generator=gen, device=device).tolist()
    y0, x0 = int(geo[0] * mask_side / 2), int(geo[1] * mask_side / 2)
    hh, ww = int(mask_side / 8 + geo[2] * mask_side * 3 / 8), int(mask_side / 8 + geo[3] * mask_side * 3 / 8)
    m = torch.zeros((mask_side, mask_side), dtype=torch.float32, device=device)
    m[y0:y0 + hh, x0:x0 + ww] = 1.0
    rim = max(mask_side // 37, 1)
    m[y0:y0 + rim, x0:x0 + ww] = 0.5
    m[y0:y0 + hh, x0:x0 + rim] = 0.25
    return feats.contiguous(), m
