"""Per-layer device timings for the five BASELINE.json configs on one GPU (CUDA events, median of N calls, inputs
resident in HBM).  Writes a markdown table (stdout).  The headline metric lives in bench.py; this is the breakdown
that profiles/ keeps per round."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from maskrcnn_tf2_b200 import functional as F
from maskrcnn_tf2_b200 import make_config, synth

dev = torch.device("cuda:0")
PEAK = 6535.4
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2] * 1e3  # us


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def align_bytes(B, N, ph, pw, S, C=256):
    maps = sum((S // s) ** 2 * C * 4 for s in (4, 8, 16, 32))
    out = N * ph * pw * C * 4
    return B * (out + min(4 * out, maps) + 16 * N)


def inference(cfg_id, B, S, NC, regime, rows, distinct=8):
    x = synth.inference_batch(cfg_id, min(B, distinct), img_size=S, num_classes=NC, regime=regime)
    rep = (B + min(B, distinct) - 1) // min(B, distinct)
    tile = lambda a: np.concatenate([a] * rep, 0)[:B]
    d = {k: t(tile(v)) for k, v in x.items() if k != "feature_maps"}
    maps = [t(tile(f)) for f in x["feature_maps"]]
    cfg = make_config(img_size=S, num_classes=NC, batch_size=B)
    prop = lambda: F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], d["anchors"], 6000, 1000, SD, 0.7)
    rois = prop()
    a7 = lambda: F.roialign_forward(rois, d["image_meta"], maps, (7, 7))
    det = lambda: F.detection_forward(rois, d["mrcnn_class"], d["mrcnn_bbox"], d["image_meta"], SD, 0.7, 100, 0.3,
                                      return_boxes=True)
    boxes = det()[1]
    a14 = lambda: F.roialign_forward(boxes, d["image_meta"], maps, (14, 14))
    tp, t7, td, t14 = timed(prop), timed(a7), timed(det), timed(a14)
    total = tp + t7 + td + t14
    name = f"cfg{cfg_id} B={B} S={S} NC={NC} {regime}"
    rows.append(f"| {name} | {tp:.0f} | {t7:.0f} ({align_bytes(B, 1000, 7, 7, S) / t7 / 1e3:.0f} GB/s closed-form) | {td:.0f} | "
                f"{t14:.0f} ({align_bytes(B, 100, 14, 14, S) / t14 / 1e3:.0f} GB/s) | {total:.0f} | {B / total * 1e6:.0f} |")
    del maps, d
    torch.cuda.empty_cache()


def training(rows, B=8, S=1024, T=200, mini=None):
    cfg_id = 3
    x = synth.inference_batch(cfg_id, B, img_size=S, regime="clustered")
    g = synth.training_targets_batch(cfg_id, B, img_size=S, mini_mask=mini)
    d = {k: t(v) for k, v in x.items() if k != "feature_maps"}
    maps = [t(f) for f in x["feature_maps"]]
    prop = F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], d["anchors"], 6000, 2000, SD, 0.7)
    gtc, gtb, gtm = t(g["gt_class_ids"]), t(g["gt_boxes"]), t(g["gt_masks"])
    keys = torch.randint(-2 ** 31, 2 ** 31, (B, 2000), device=dev, dtype=torch.int64).to(torch.int32)
    tgt = lambda: F.detection_target_forward(prop, gtc, gtb, gtm, keys, T, 0.33, SD, (28, 28),
                                             use_mini_masks=mini is not None, return_counts=True)
    rois, _, _, _, counts = tgt()
    shapes = [tuple(m.shape) for m in maps]
    res = {}
    for ph in (7, 14):
        out, roi_map = F.roialign_forward(rois, d["image_meta"], maps, (ph, ph))
        gout = torch.randn_like(out)
        res[ph] = (timed(lambda: F.roialign_forward(rois, d["image_meta"], maps, (ph, ph))),
                   timed(lambda: F.roialign_backward(gout, rois, roi_map, shapes, deterministic=False)),
                   timed(lambda: F.roialign_backward(gout, rois, roi_map, shapes, deterministic=True)))
    tt = timed(tgt)
    mb = sum(np.prod(s) for s in shapes) * 4
    name = f"cfg3 B={B} S={S} T={T} masks={'mini32' if mini else 'full'} (pos/neg {counts.float().mean(0).tolist()})"
    rows.append(f"| {name} | DetectionTarget {tt:.0f} | 7x7 fwd {res[7][0]:.0f} / bwd atomic {res[7][1]:.0f} "
                f"({(B * T * 49 * 1024 + mb) / res[7][1] / 1e3:.0f} GB/s), deterministic {res[7][2]:.0f} | 14x14 fwd "
                f"{res[14][0]:.0f} / bwd atomic {res[14][1]:.0f} ({(B * T * 196 * 1024 + mb) / res[14][1] / 1e3:.0f} GB/s), "
                f"deterministic {res[14][2]:.0f} | | | |")


if __name__ == "__main__":
    rows = []
    inference(1, 1, 1024, 2, "clustered", rows)
    inference(2, 8, 1024, 81, "clustered", rows)
    inference(2, 8, 1024, 81, "sparse", rows)
    inference(2, 8, 1024, 81, "iid", rows)
    inference(4, 32, 512, 81, "clustered", rows)
    inference(4, 32, 256, 81, "clustered", rows)
    inference(5, 16, 1024, 81, "clustered", rows)
    inference(5, 32, 1024, 81, "clustered", rows)
    inference(5, 64, 1024, 81, "clustered", rows)
    print("| inference config | Proposal us | ROIAlign 7x7 us | Detection us | ROIAlign 14x14 us | stage us | images/s |")
    print("|---|---|---|---|---|---|---|")
    print("\n".join(rows))
    rows = []
    training(rows)
    training(rows, mini=(32, 32))
    training(rows, T=2000)
    print()
    print("| training config | DetectionTarget us | ROIAlign 7x7 | ROIAlign 14x14 | | | |")
    print("|---|---|---|---|---|---|---|")
    print("\n".join(rows))
