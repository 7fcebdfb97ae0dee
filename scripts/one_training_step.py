"""One training-side pass at config 3 (B=8, COCO shape, T=200): ProposalLayer (2000) -> DetectionTargetLayer ->
PyramidROIAlign 7x7 forward + backward (atomic and deterministic) + Proposal gradient, for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
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
keys = torch.randint(-2 ** 31, 2 ** 31, (B, 2000), device=dev, dtype=torch.int64).to(torch.int32)
for _ in range(2):
    r = F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], d["anchors"], 6000, 2000, SD, 0.7, debug=True)
    rois = F.detection_target_forward(r["proposals"], t(g["gt_class_ids"]), t(g["gt_boxes"]), t(g["gt_masks"]), keys, T,
                                      0.33, SD, (28, 28))[0]
    out, roi_map = F.roialign_forward(rois, d["image_meta"], maps, (7, 7))
    go = torch.ones_like(out)
    ga = F.roialign_backward(go, rois, roi_map, shapes, deterministic=False)
    gd = F.roialign_backward(go, rois, roi_map, shapes, deterministic=True)
    gp = F.proposal_backward(torch.ones_like(r["proposals"]), d["rpn_bbox"], d["anchors"], r["topk_idx"], r["keep_idx"], SD)
torch.cuda.synchronize()
print("ok", float(out.sum()), float(ga[0].sum()), float(gd[0].sum()), float(gp.sum()))
