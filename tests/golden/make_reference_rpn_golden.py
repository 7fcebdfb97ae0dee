"""Generates tests/golden/reference_rpn_targets_golden.npz by RUNNING THE REFERENCE'S OWN `build_rpn_targets`
(utils.py:154-262, plain numpy; called per image by the data loader, preprocess.py:342-348) in this container
(`python tests/golden/make_reference_rpn_golden.py`; needs /root/reference, so it cannot run on the GPU box -- the
vectors it writes are committed).

The one non-deterministic step of that function, `np.random.choice(ids, extra, replace=False)` (utils.py:219,227),
is replaced WHILE THE REFERENCE RUNS by a key-driven stand-in with the rule the CUDA path and the oracle implement:
of the candidate ids keep the `len(ids) - extra` with the largest injected key (ties -> lower anchor index) and
return the others.  Everything else -- float64 IoU, crowd handling, the `overlaps == max` matching, the order of
the rpn_bbox rows, the float64 box-refinement arithmetic -- is the reference's code, untouched.

Nothing is copied from the reference: the module is loaded from where it lies and only its outputs are stored.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_reference_golden import load_reference_utils, reference_config  # noqa: E402

RPN_BBOX_STD = np.array([0.1, 0.1, 0.2, 0.2])  # src/common/config.py:90


def keyed_choice(keys):
    def choice(ids, extra, replace=False):
        assert replace is False
        ids = np.asarray(ids)
        order = np.lexsort((ids, -keys[ids].astype(np.float64)))  # key desc, index asc
        return ids[order[len(ids) - int(extra):]]
    return choice


def run_reference(U, anchors, class_ids, boxes, R, keys):
    """One image through the reference's build_rpn_targets with the keyed stand-in for np.random.choice."""
    saved = np.random.choice
    np.random.choice = keyed_choice(keys)
    try:
        real = class_ids != 0                                        # the loader passes only real instances
        match, bbox = U.build_rpn_targets(anchors, class_ids[real], boxes[real], R, RPN_BBOX_STD)
    finally:
        np.random.choice = saved
    return match, bbox


def pixel_anchors(U, S):
    cfg = reference_config(S)
    return U.generate_pyramid_anchors(cfg["rpn_anchor_scales"], cfg["rpn_anchor_ratios"],
                                      U.compute_backbone_shapes(cfg), cfg["backbone_strides"], cfg["rpn_anchor_stride"])


def small_cases(S=128, G=8):
    """[B,G] class ids (0 = padding row, negative = crowd) and [B,G,4] int32 pixel boxes."""
    cls = np.zeros((6, G), np.int32)
    box = np.zeros((6, G, 4), np.int32)
    # 0: five ordinary instances
    box[0, :5] = [[10, 12, 60, 70], [30, 40, 100, 120], [64, 0, 128, 50], [5, 90, 40, 126], [80, 70, 112, 102]]
    cls[0, :5] = [3, 1, 7, 2, 2]
    # 1: four instances and one crowd region
    box[1, :5] = [[0, 0, 64, 64], [20, 70, 52, 102], [70, 20, 118, 44], [90, 90, 122, 122], [40, 40, 128, 128]]
    cls[1, :5] = [1, 2, 3, 4, -1]
    # 2: instances that coincide with anchors (IoU exactly 1, many exact ties) -> more positives than R/2
    box[2, :6] = [[0, 0, 32, 32], [16, 16, 48, 48], [32, 32, 96, 96], [0, 64, 64, 128], [48, 0, 80, 32], [64, 64, 96, 96]]
    cls[2, :6] = [1, 1, 2, 2, 3, 3]
    # 3: a zero-area instance: its IoU column is all zeros, `overlaps == max` marks EVERY anchor positive (L:208-209)
    box[3, :2] = [[20, 20, 90, 100], [50, 50, 50, 80]]
    cls[3, :2] = [5, 6]
    # 4: one tiny instance no anchor reaches 0.3 with: only the forced best match is positive
    box[4, :1] = [[61, 67, 66, 71]]
    cls[4, :1] = [9]
    # 5: padding rows between real ones, two crowds
    box[5, [0, 2, 3, 6, 7]] = [[8, 8, 72, 40], [30, 60, 94, 124], [0, 0, 128, 128], [100, 10, 124, 58], [60, 60, 128, 128]]
    cls[5, [0, 2, 3, 6, 7]] = [2, 4, -1, 6, -3]
    return cls, box


def coco_cases(B=2, S=1024, G=100, n_real=20, seed=4200):
    cls = np.zeros((B, G), np.int32)
    box = np.zeros((B, G, 4), np.int32)
    for b in range(B):
        rng = np.random.default_rng(seed + b)
        side = rng.uniform(32, 512, (n_real, 2))
        c = rng.uniform(0, S, (n_real, 2))
        y1 = np.clip(c[:, 0] - side[:, 0] / 2, 0, S - 2)
        x1 = np.clip(c[:, 1] - side[:, 1] / 2, 0, S - 2)
        y2 = np.clip(c[:, 0] + side[:, 0] / 2, y1 + 2, S)
        x2 = np.clip(c[:, 1] + side[:, 1] / 2, x1 + 2, S)
        box[b, :n_real] = np.round(np.stack([y1, x1, y2, x2], 1)).astype(np.int32)
        cls[b, :n_real] = rng.integers(1, 81, n_real)
        if b == 1:
            cls[b, 3] = -1                                           # one crowd
    return cls, box


def keys_for(seed, b, A):
    return np.random.default_rng(seed * 1000 + b).random(A, dtype=np.float32)


def build():
    U = load_reference_utils()
    g = {}
    # ---- small: everything stored --------------------------------------------------------------------
    S, R = 128, 64
    anchors = pixel_anchors(U, S)
    A = anchors.shape[0]
    cls, box = small_cases(S)
    B = cls.shape[0]
    match = np.zeros((B, A), np.int32)
    bbox = np.zeros((B, R, 4), np.float64)
    match_all = np.zeros((B, A), np.int32)                           # no subsampling (R = 2A): the matching rule alone
    for b in range(B):
        k = keys_for(7, b, A)
        match[b], bbox[b] = run_reference(U, anchors, cls[b], box[b], R, k)
        match_all[b], _ = run_reference(U, anchors, cls[b], box[b], 2 * A, k)
    g.update(small_S=np.int64(S), small_R=np.int64(R), small_key_seed=np.int64(7), small_gt_class_ids=cls,
             small_gt_boxes=box, small_rpn_match=match.astype(np.int8), small_rpn_bbox=bbox,
             small_rpn_match_unsampled=match_all.astype(np.int8))
    # ---- COCO shape (A = 261888): inputs + index lists + digest ------------------------------------------
    S, R = 1024, 256
    anchors = pixel_anchors(U, S)
    A = anchors.shape[0]
    cls, box = coco_cases()
    B = cls.shape[0]
    pos = -np.ones((B, R), np.int32)
    neg = -np.ones((B, R), np.int32)
    bbox = np.zeros((B, R, 4), np.float64)
    dig = np.zeros((B, 32), np.uint8)
    for b in range(B):
        m, bbox[b] = run_reference(U, anchors, cls[b], box[b], R, keys_for(11, b, A))
        p, n = np.where(m == 1)[0], np.where(m == -1)[0]
        assert len(p) + len(n) <= R
        pos[b, :len(p)], neg[b, :len(n)] = p, n
        dig[b] = np.frombuffer(hashlib.sha256(m.astype(np.int32).tobytes()).digest(), np.uint8)
    g.update(coco_S=np.int64(S), coco_R=np.int64(R), coco_key_seed=np.int64(11), coco_gt_class_ids=cls,
             coco_gt_boxes=box, coco_pos_idx=pos, coco_neg_idx=neg, coco_rpn_bbox=bbox, coco_rpn_match_sha256=dig)
    return g


if __name__ == "__main__":
    out = os.path.join(HERE, "reference_rpn_targets_golden.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, os.path.getsize(out), "bytes")
    z = np.load(out)
    for b in range(z["small_rpn_match"].shape[0]):
        m, mu = z["small_rpn_match"][b], z["small_rpn_match_unsampled"][b]
        print(f"small image {b}: +{(m == 1).sum()} -{(m == -1).sum()}  (unsampled +{(mu == 1).sum()} -{(mu == -1).sum()})")
    for b in range(z["coco_pos_idx"].shape[0]):
        print(f"coco image {b}: +{(z['coco_pos_idx'][b] >= 0).sum()} -{(z['coco_neg_idx'][b] >= 0).sum()}")
