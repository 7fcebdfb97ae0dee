"""Generates tests/golden/roi_stage_golden.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference itself cannot run here (no TensorFlow), so these vectors pin
the oracle's behaviour over time and give the CUDA path a fixed target that does not depend on rebuilding the oracle.
Inputs are seeded; every array needed to replay a case is stored next to its expected outputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from conftest import random_boxes  # noqa: E402
from maskrcnn_tf2_b200 import synth  # noqa: E402

SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)


def build():
    g = {}
    rng = np.random.default_rng(20261018)
    # top-k with ties
    s = (np.round(rng.uniform(0, 1, (2, 3000)) * 64) / 64).astype(np.float32)
    g["topk_scores"] = s
    g["topk_idx"] = np.stack([oracle.topk(s[b], 500) for b in range(2)])
    # NMS (clustered boxes, tied scores, degenerate rows)
    boxes = np.stack([random_boxes(rng, 600, clusters=8) for _ in range(2)])
    scores = (np.round(rng.uniform(0, 1, (2, 600)) * 32) / 32).astype(np.float32)
    boxes[:, 50:60, 2] = boxes[:, 50:60, 0]
    scores[0, 5] = -np.inf
    g["nms_boxes"], g["nms_scores"] = boxes, scores
    keep = np.full((2, 100), -1, np.int32)
    cnt = np.zeros(2, np.int32)
    for b in range(2):
        k = oracle.nms(boxes[b], scores[b], 100, 0.5)
        keep[b, :len(k)] = k
        cnt[b] = len(k)
    g["nms_keep"], g["nms_count"] = keep, cnt
    # ProposalLayer on a 128-px anchor pyramid
    a = synth.pyramid_anchors(128)
    pr, bb = zip(*[synth.rpn_outputs(np.random.default_rng(7 + b), a, "sparse", 128) for b in range(2)])
    pr, bb, an = np.stack(pr), np.stack(bb), np.ascontiguousarray(np.broadcast_to(a, (2,) + a.shape))
    r = oracle.proposal_layer(pr, bb, an, 1000, 200, SD, 0.7)
    g["prop_probs"], g["prop_bbox"], g["prop_anchors"] = pr, bb, an
    g["prop_out"], g["prop_topk"], g["prop_keep"], g["prop_count"] = (r["proposals"], r["topk_idx"], r["keep_idx"],
                                                                      r["keep_count"])
    # PyramidROIAlign forward + gradient (generic channel count), zero-padded and out-of-range ROIs
    rois = r["proposals"][:, :24].copy()
    rois[:, -3:] = 0.0
    rois[0, 0] = [-0.2, 0.1, 0.4, 1.3]
    fm = [rng.standard_normal((2, h, h, 12)).astype(np.float32) for h in (32, 16, 8, 4)]
    ra = oracle.pyramid_roi_align(rois, 128.0, 128.0, fm, (7, 7))
    grad = rng.standard_normal(ra["out"].shape).astype(np.float32)
    rg = oracle.pyramid_roi_align_grad(grad, rois, 128.0, 128.0, [f.shape for f in fm])
    g["ra_boxes"], g["ra_grad"] = rois, grad
    for l in range(4):
        g[f"ra_fm{l}"], g[f"ra_gfm{l}"] = fm[l], rg[l]
    g["ra_out"], g["ra_map"], g["ra_level"] = ra["out"], ra["roi_map"], ra["level"]
    # DetectionLayer
    drois = np.stack([random_boxes(rng, 80, clusters=5) for _ in range(2)])
    probs, deltas = synth.head_outputs(rng, 2, 80, 6)
    meta = synth.image_meta(2, 128, 6)
    meta[1, 7:11] = (8, 16, 120, 112)
    d = oracle.detection_layer(drois, probs, deltas, meta, SD, 0.5, 20, 0.3)
    g["det_rois"], g["det_probs"], g["det_deltas"], g["det_meta"] = drois, probs, deltas, meta
    g["det_out"], g["det_count"] = d["detections"], d["count"]
    # DetectionTargetLayer
    props = np.stack([random_boxes(rng, 120, min_size=0.05, max_size=0.5, clusters=6) for _ in range(2)])
    props[:, -10:] = 0.0
    gtb = np.zeros((2, 8, 4), np.float32)
    gtc = np.zeros((2, 8), np.int32)
    for b in range(2):
        gtb[b, :5] = props[b, rng.choice(110, 5, replace=False)]
        gtc[b, :5] = rng.integers(1, 6, 5)
    gtc[1, 4] = -2
    masks = (rng.uniform(0, 1, (2, 24, 24, 8)) < 0.5).astype(np.uint8)
    keys = rng.integers(0, 2 ** 32, (2, 120), dtype=np.uint64).astype(np.uint32)
    t = oracle.detection_target_layer(props, gtc, gtb, masks, keys, 32, 0.33, SD, (28, 28))
    g["dt_props"], g["dt_gtc"], g["dt_gtb"], g["dt_masks"], g["dt_keys"] = props, gtc, gtb, masks, keys
    g["dt_rois"], g["dt_cls"], g["dt_deltas"], g["dt_out_masks"], g["dt_counts"] = (t["rois"], t["class_ids"],
                                                                                      t["deltas"], t["masks"],
                                                                                      t["counts"])
    return g


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "roi_stage_golden.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, os.path.getsize(out), "bytes")
