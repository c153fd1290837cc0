"""How many same-label pairs with overlapping full-resolution boxes does an image of the benchmark produce?
(work statistics of compute_semantic_ios on the synthetic workload; torch port on the GPU)"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_torch
pkg = importlib.import_module("no-time-to-train_b200")
synth = pkg.synth
dev = torch.device("cuda", 0)
c = 1024
centres = synth.cluster_centres(c)
gen = torch.Generator().manual_seed(7)
bank = centres[torch.arange(80) % 5].unsqueeze(1) + (0.3 / c ** 0.5) * torch.randn(80, 10, c, generator=gen)
for seed in (1234, 1235):
    lr, iou, feat = synth.make_stage_inputs_device(1024, centres, dev, seed=seed)
    with torch.inference_mode():
        out = ref_torch.match_image(lr, iou, feat, bank.to(dev), ref_torch.StageConfig(num_out_instance=100), (1024, 1024))
    aux = out["aux"]
    labels = aux["labels_all"][aux["sel_index"]]
    b = aux["full_boxes"].float()
    k = labels.numel()
    same = labels[:, None] == labels[None, :]
    ov = (torch.maximum(b[:, None, 0], b[None, :, 0]) <= torch.minimum(b[:, None, 2], b[None, :, 2])) & \
         (torch.maximum(b[:, None, 1], b[None, :, 1]) <= torch.minimum(b[:, None, 3], b[None, :, 3]))
    up = torch.triu(torch.ones(k, k, dtype=torch.bool, device=dev), 1)
    pairs = same & ov & up
    x0 = torch.maximum(b[:, None, 0], b[None, :, 0]); x1 = torch.minimum(b[:, None, 2], b[None, :, 2])
    y0 = torch.maximum(b[:, None, 1], b[None, :, 1]); y1 = torch.minimum(b[:, None, 3], b[None, :, 3])
    words = (((x1 / 32).floor() - (x0 / 32).floor() + 1) * (y1 - y0 + 1))[pairs]
    print(f"seed {seed}: K={k} labels used={labels.unique().numel()} pairs={int(pairs.sum())} "
          f"window words: total={int(words.sum())} mean={float(words.mean()):.0f} max={int(words.max())} "
          f">256: {int((words > 256).sum())} >4096: {int((words > 4096).sum())}")
    print("  label histogram top:", torch.bincount(labels).sort(descending=True).values[:8].tolist())
