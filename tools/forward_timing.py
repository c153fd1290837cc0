"""Where does one bs=1 call of the drop-in stage spend its time on the host?  (perf_counter around each piece)"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("no-time-to-train_b200")
synth = pkg.synth
dev = torch.device("cuda", 0)
c = 1024
centres = synth.cluster_centres(c)
gen = torch.Generator().manual_seed(7)
bank = centres[torch.arange(80) % 5].unsqueeze(1) + (0.3 / c ** 0.5) * torch.randn(80, 10, c, generator=gen)
stage = pkg.MatchingStage(dev, pkg.StageConfig(nms_thr=0.5, num_out_instance=100, enc_hw=(37, 37)))
stage.set_prototypes(bank)
images = [synth.make_stage_inputs_device(1024, centres, dev, seed=1234 + i) for i in range(4)]
outs = (torch.zeros((100, 1024, 1024), dtype=torch.uint8, device=dev), torch.zeros((100, 4), dtype=torch.int32, device=dev))
acc = {}
_real = stage.lib.nttt_match_image
def _timed(*a):
    t0 = time.perf_counter()
    r = _real(*a)
    acc["  nttt_match_image (C call)"] = acc.get("  nttt_match_image (C call)", 0.0) + time.perf_counter() - t0
    return r
class _Lib:
    def __getattr__(self, k):
        return _timed if k == "nttt_match_image" else getattr(_real_lib, k)
_real_lib = stage.lib
stage.lib = _Lib()
def lap(name, t0):
    torch.cuda.synchronize(); t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + t - t0; return t
for rep in range(44):
    if rep == 4:
        acc.clear()  # the first reps pay the one-time costs (lazy kernel loading, tables, workspaces)
    img = images[rep % 4]
    torch.cuda.synchronize(); t = time.perf_counter()
    for nth, rle in enumerate((False, True, False, True) if rep % 2 else (True, False, True, False)):
        tag = ("rle" if rle else "plain") + (" 1st" if nth < 2 else " 2nd") + (" (first call of the rep)" if nth == 0 else "")
        p = stage.match_async(*img, (1024, 1024), slot=0, persistent_out=outs, rle=rle, low_latency=True)
        t1 = time.perf_counter(); acc[tag + " enqueue (host)"] = acc.get(tag + " enqueue (host)", 0.0) + t1 - t
        t = lap(tag + " enqueue+kernels", t)
        out = p.get()
        t = lap(tag + " get()", t)
        if rle:
            segs = p.rle_segmentations()
            t = lap("rle_segmentations()", t)
        del p, out
for k, v in acc.items():
    print(f"{k:48s} {1e6 * v / 40:8.1f} us per rep (divide the per-variant rows by their share of reps)")
