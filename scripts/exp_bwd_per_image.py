"""Experiment: does the atomic ROIAlign backward get faster when each image's maps are zeroed right before that
image's ROIs are scattered (zeroed lines still in L2)?  Compares one batched call with B per-image calls."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maskrcnn_tf2_b200 import functional as F, synth
dev = torch.device("cuda:0")
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
B, S, T = 8, 1024, 200
x = synth.inference_batch(3, B, img_size=S, regime="clustered")
g = synth.training_targets_batch(3, B, img_size=S)
d = {k: t(v) for k, v in x.items() if k != "feature_maps"}
maps = [t(f) for f in x["feature_maps"]]
shapes = [tuple(m.shape) for m in maps]
shapes1 = [(1,) + s[1:] for s in shapes]
prop = F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], d["anchors"], 6000, 2000, SD, 0.7)
keys = torch.randint(-2 ** 31, 2 ** 31, (B, 2000), device=dev, dtype=torch.int64).to(torch.int32)
rois = F.detection_target_forward(prop, t(g["gt_class_ids"]), t(g["gt_boxes"]), t(g["gt_masks"]), keys, T, 0.33, SD, (28, 28))[0]
def timed(fn, reps=10):
    for _ in range(2): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2] * 1e3
for ph in (7, 14):
    out, roi_map = F.roialign_forward(rois, d["image_meta"], maps, (ph, ph), map_mode=1)
    go = torch.randn_like(out)
    whole = timed(lambda: F.roialign_backward(go, rois, roi_map, shapes, deterministic=False))
    def per_image():
        for b in range(B):
            F.roialign_backward(go[b:b + 1], rois[b:b + 1], roi_map[b:b + 1], shapes1, deterministic=False)
    print(f"{ph}x{ph}: batched {whole:.0f} us, per image x{B}: {timed(per_image):.0f} us")
