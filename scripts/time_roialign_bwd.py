"""Times PyramidROIAlign backward at the training shapes (config 3): deterministic gather vs atomic scatter.
CUDA events, median of 20, maps larger than L2.  python scripts/time_roialign_bwd.py [T ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maskrcnn_tf2_b200 import functional as F  # noqa: E402
from maskrcnn_tf2_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2] * 1e3


def main():
    B, S = 8, 1024
    Ts = [int(a) for a in sys.argv[1:]] or [200, 2000]
    x = synth.inference_batch(3, B, img_size=S, regime="clustered")
    g = synth.training_targets_batch(3, B, img_size=S, mini_mask=None)
    d = {k: t(v) for k, v in x.items() if k != "feature_maps"}
    maps = [t(f) for f in x["feature_maps"]]
    shapes = [tuple(m.shape) for m in maps]
    mb = sum(int(np.prod(s)) for s in shapes) * 4
    prop = F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], d["anchors"], 6000, 2000, SD, 0.7)
    keys = torch.randint(-2 ** 31, 2 ** 31, (B, 2000), device=dev, dtype=torch.int64).to(torch.int32)
    for T in Ts:
        rois, _, _, _, counts = F.detection_target_forward(prop, t(g["gt_class_ids"]), t(g["gt_boxes"]), t(g["gt_masks"]),
                                                           keys, T, 0.33, SD, (28, 28), return_counts=True)
        for ph in (7, 14):
            out, roi_map = F.roialign_forward(rois, d["image_meta"], maps, (ph, ph))
            gout = torch.randn_like(out)
            alg = B * T * ph * ph * 1024 + mb
            res = {}
            for det in (True, False):
                us = timed(lambda: F.roialign_backward(gout, rois, roi_map, shapes, deterministic=det))
                res[det] = us
            a = F.roialign_backward(gout, rois, roi_map, shapes, deterministic=True)
            b = F.roialign_backward(gout, rois, roi_map, shapes, deterministic=False)
            err = max((p - q).abs().max().item() for p, q in zip(a, b))
            print(f"T={T} {ph}x{ph} real rois/img {counts.sum(1).float().mean().item():.0f}: gather {res[True]:.0f} us "
                  f"({alg / res[True] / 1e3:.0f} GB/s)  atomic {res[False]:.0f} us ({alg / res[False] / 1e3:.0f} GB/s)  "
                  f"algorithmic {alg / 1e6:.0f} MB  max|gather-atomic| {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
