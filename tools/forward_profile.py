"""cProfile of the drop-in class's bs=1 forward loop (host side): where the 0.6 ms per image go."""
import cProfile, importlib, os, pstats, sys, io
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
resident = [synth.make_stage_inputs_device(1024, centres, dev, seed=1234 + i) for i in range(8)]

class SeamModel(pkg.Sam2MatchingBaselineNoAMG):
    def _extract_target_features(self, tar_img, device):
        return resident[self._cur][2], tar_img
    def _forward_sam(self, imgs):
        lr, iou, _ = resident[self._cur]
        return lr, iou, None

class Dataset:
    def __len__(self):
        return 256
    def __getitem__(self, i):
        model._cur = i % 8
        return dict(target_img=torch.zeros(3, 8, 8), target_img_info=dict(ori_height=1024, ori_width=1024, file_name=f"s{i}", id=i))

model = SeamModel(sam2_infer_cfgs=dict(points_per_side=32, testing_point_bs=256, iou_thr=0.0, nms_thr=0.5, num_out_instance=100,
                                       kmeans_k=2, n_pca_components=2, cls_num_per_mask=1),
                  memory_bank_cfg=dict(enable=True, category_num=80, length=10), encoder_geometry=(518, 14, c), device=dev)
model.memory_bank.feats_ins_avg.copy_(bank.to(dev)); model.memory_bank.postprocessed[0] = True; model.memory_bank.ready = True
runner = pkg.MatcherRunner(model, "test", Dataset(), rle=True)
runner.run()  # warm-up pass
runner.setup()
pr = cProfile.Profile()
pr.enable()
runner.run()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(38)
print(s.getvalue()[:6000])
print("forward ms/image", 1e3 * sum(runner.time_queue) / len(runner.time_queue))
