"""Per-layer device times of one configuration on one GPU (CUDA events, median of N eager calls, inputs resident in
HBM, an L2-sized buffer written between calls).  python scripts/layer_times.py [--batch 8] [--img 1024] [--train]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from maskrcnn_tf2_b200 import functional as F
from maskrcnn_tf2_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--img", type=int, default=1024)
ap.add_argument("--nc", type=int, default=81)
ap.add_argument("--cfg", type=int, default=2)
ap.add_argument("--train", action="store_true")
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--tag", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=a.iters, warm=3):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for x, y in ev:
        flush.zero_()
        x.record(); fn(); y.record()
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    return t[len(t) // 2] * 1e3


def t(v):
    return torch.from_numpy(np.ascontiguousarray(v)).to(dev)


B, S, NC = a.batch, a.img, a.nc
x = synth.inference_batch(a.cfg, B, img_size=S, num_classes=NC, regime="clustered")
d = {k: t(v) for k, v in x.items() if k != "feature_maps"}
maps = [t(f) for f in x["feature_maps"]]
shapes = [tuple(m.shape) for m in maps]
res = {}
if not a.train:
    tk = lambda: F.topk(d["rpn_probs"], 6000, column=1)
    prop = lambda: F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], d["anchors"], 6000, 1000, SD, 0.7)
    rois = prop()
    a7 = lambda: F.roialign_forward(rois, d["image_meta"], maps, (7, 7))
    det = lambda: F.detection_forward(rois, d["mrcnn_class"], d["mrcnn_bbox"], d["image_meta"], SD, 0.7, 100, 0.3,
                                      return_boxes=True)
    boxes = det()[1]
    a14 = lambda: F.roialign_forward(boxes, d["image_meta"], maps, (14, 14))
    res = {"topk": timed(tk), "proposal": timed(prop), "align7": timed(a7), "detection": timed(det), "align14": timed(a14)}
    res["sum"] = res["proposal"] + res["align7"] + res["detection"] + res["align14"]
else:
    g = synth.training_targets_batch(a.cfg, B, img_size=S)
    keys = np.random.default_rng(9).integers(0, 2 ** 32, (B, 2000), dtype=np.uint64).astype(np.uint32)
    prop = F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], d["anchors"], 6000, 2000, SD, 0.7)
    gc, gb, gm, kk = t(g["gt_class_ids"]), t(g["gt_boxes"]), t(g["gt_masks"]), t(keys.view(np.int32))
    dt = lambda: F.detection_target_forward(prop, gc, gb, gm, kk, 200, 0.33, SD, (28, 28))
    rois = dt()[0]
    res["targets"] = timed(dt)
    for pool in ((7, 7), (14, 14)):
        fw = lambda: F.roialign_forward(rois, d["image_meta"], maps, pool)
        out, rmap = fw()
        go = torch.randn_like(out)
        res[f"fwd{pool[0]}"] = timed(fw)
        res[f"bwd{pool[0]}_det"] = timed(lambda: F.roialign_backward(go, rois, rmap, shapes, deterministic=True))
        res[f"bwd{pool[0]}_atomic"] = timed(lambda: F.roialign_backward(go, rois, rmap, shapes, deterministic=False))
        del out, go
print(a.tag, f"B={B} S={S}", " ".join(f"{k}={v:.1f}us" for k, v in res.items()), flush=True)
