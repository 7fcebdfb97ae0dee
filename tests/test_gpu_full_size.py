"""Parity at BASELINE.json's FULL sizes for the configurations the bench line is not quoted on.

configs[2] (training step): 1024^2, 8 images, 2000 proposals -> DetectionTargetLayer (T = 200, full 1024^2 masks and
    the 32x32 mini-mask variant) -> PyramidROIAlign 7x7 / 14x14 forward and both backward modes.  Compared with the
    CPU oracle AND with SHA-256 digests of the outputs of the reference's own layer code on the same inputs
    (tests/golden/make_reference_layers_golden.py::build_full_size_training_config3; mrcnn_layers.py:844-1007).
stress row: 2000 ROIs per image straight through ROIAlign forward / backward.
configs[4] (batch 64): the whole batch on one GPU in the reference's first-appearance map mode (Q2, map_mode=0)
    against the oracle, and the 32 / 16 / 8-image shards a 2 / 4 / 8-GPU split computes, each against the oracle run
    on that shard and against the reference-executed digests (the Q2 table is local to a replica).
NMS at the threshold: pairs whose IoU equals the threshold exactly or sits 1-2 ulp on either side of it, placed in the
    same tile, in neighbouring tiles and tiles apart of a 6000-candidate problem (the cluster path of nms_lazy_kernel),
    for thr in {0.3, 0.5, 0.7} (tf.image.non_max_suppression, mrcnn_layers.py:225,455: strict '>', true division).
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
HERE = os.path.dirname(os.path.abspath(__file__))


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digests():
    return json.load(open(os.path.join(HERE, "golden", "reference_layers_full_size_sha256.json")))


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mini", [False, True])
def test_config3_full_size_training_step(orc, dev, mini):
    from maskrcnn_tf2_b200 import functional as F
    from maskrcnn_tf2_b200 import synth
    B, S, Tn = 8, 1024, 200
    rec = digests()["training_full_mini" if mini else "training_full"]
    x = synth.inference_batch(3, B, img_size=S, regime="clustered")
    g = synth.training_targets_batch(3, B, img_size=S, mini_mask=(32, 32) if mini else None)
    keys = np.random.default_rng(9).integers(0, 2 ** 32, (B, 2000), dtype=np.uint64).astype(np.uint32)
    for k, v in (("rpn_probs", x["rpn_probs"]), ("gt_masks", g["gt_masks"]), ("keys", keys)):
        assert sha(v) == rec["input_sha256"][k], k            # same inputs as the reference-executed run
    ref_p = orc.proposal_layer(x["rpn_probs"], x["rpn_bbox"], x["anchors"], 6000, 2000, SD, 0.7)["proposals"]
    prop = F.proposal_forward(T(x["rpn_probs"], dev), T(x["rpn_bbox"], dev), T(x["anchors"], dev), 6000, 2000, SD, 0.7)
    assert np.array_equal(N(prop), ref_p)
    assert sha(N(prop)) == rec["sha256"]["proposals"]
    ref_t = orc.detection_target_layer(ref_p, g["gt_class_ids"], g["gt_boxes"], g["gt_masks"], keys, Tn, 0.33, SD,
                                       (28, 28), use_mini_masks=mini)
    rois, cls, dl, mk = F.detection_target_forward(prop, T(g["gt_class_ids"], dev), T(g["gt_boxes"], dev),
                                                   T(g["gt_masks"], dev), T(keys.view(np.int32), dev), Tn, 0.33, SD,
                                                   (28, 28), use_mini_masks=mini)
    for got, name in [(rois, "rois"), (cls, "class_ids"), (dl, "deltas"), (mk, "masks")]:
        assert np.array_equal(N(got), ref_t[name]), name
        assert sha(N(got)) == rec["sha256"][name], name      # = the reference's own DetectionTargetLayer output
    assert [int((N(cls)[b] != 0).sum()) for b in range(B)] == rec["positives"]
    if mini:
        return
    fm = [T(f, dev) for f in x["feature_maps"]]
    shapes = [tuple(f.shape) for f in x["feature_maps"]]
    meta = T(x["image_meta"], dev)
    rng = np.random.default_rng(10)
    for pool, name in (((7, 7), "pooled7"), ((14, 14), "pooled14")):
        ref_f = orc.pyramid_roi_align(ref_t["rois"], float(S), float(S), x["feature_maps"], pool)
        out, roi_map = F.roialign_forward(rois, meta, fm, pool)
        assert np.array_equal(N(roi_map), ref_f["roi_map"])
        assert np.array_equal(N(out), ref_f["out"])
        assert sha(N(out)) == rec["sha256"][name]            # = the reference's own PyramidROIAlign output
        go = rng.standard_normal(ref_f["out"].shape).astype(np.float32)
        ref_b = orc.pyramid_roi_align_grad(go, ref_t["rois"], float(S), float(S), shapes)
        tgo = T(go, dev)
        det = F.roialign_backward(tgo, rois, roi_map, shapes, deterministic=True)
        det2 = F.roialign_backward(tgo, rois, roi_map, shapes, deterministic=True)
        for l in range(4):
            assert torch.equal(det[l], det2[l])              # reproducible
            assert np.array_equal(N(det[l]), ref_b[l]), (pool, l)   # TF's sequential accumulation order, bit for bit
        del det2
        atm = F.roialign_backward(tgo, rois, roi_map, shapes, deterministic=False)
        mag = orc.pyramid_roi_align_grad(np.abs(go), ref_t["rois"], float(S), float(S), shapes)
        for l in range(4):
            # unordered fp32 accumulation: the error bound is relative to the sum of magnitudes (DESIGN.md Numerics)
            tol = 1e-6 + 1e-5 * np.maximum(np.abs(ref_b[l]), mag[l])
            assert np.all(np.abs(N(atm[l]) - ref_b[l]) <= tol), (pool, l)
        del det, atm, out, tgo
        torch.cuda.empty_cache()


@pytest.mark.parametrize("pool,B", [((7, 7), 8), ((14, 14), 4)])
def test_stress_2000_rois_per_image_through_roialign(orc, dev, pool, B):
    """BASELINE.json configs[2] says "2000 RoIs/image": besides the T = 200 reading (above) the 2000 proposals themselves
    go through PyramidROIAlign forward and backward (SURVEY 8(d) "stress row")."""
    from maskrcnn_tf2_b200 import functional as F
    from maskrcnn_tf2_b200 import synth
    S = 1024
    x = synth.inference_batch(3, B, img_size=S, regime="clustered")
    prop = F.proposal_forward(T(x["rpn_probs"], dev), T(x["rpn_bbox"], dev), T(x["anchors"], dev), 6000, 2000, SD, 0.7)
    rois = N(prop)
    assert all(int(rois[b].any(-1).sum()) > 1000 for b in range(B))
    fm = [T(f, dev) for f in x["feature_maps"]]
    shapes = [tuple(f.shape) for f in x["feature_maps"]]
    ref_f = orc.pyramid_roi_align(rois, float(S), float(S), x["feature_maps"], pool)
    out, roi_map = F.roialign_forward(prop, T(x["image_meta"], dev), fm, pool)
    assert np.array_equal(N(roi_map), ref_f["roi_map"])
    assert np.array_equal(N(out), ref_f["out"])
    del out
    go = np.random.default_rng(11).standard_normal(ref_f["out"].shape, dtype=np.float32)
    del ref_f
    ref_b = orc.pyramid_roi_align_grad(go, rois, float(S), float(S), shapes)
    mag = orc.pyramid_roi_align_grad(np.abs(go), rois, float(S), float(S), shapes)
    tgo = T(go, dev)
    for deterministic in (True, False):
        got = F.roialign_backward(tgo, prop, roi_map, shapes, deterministic=deterministic)
        for l in range(4):
            tol = 1e-6 + 1e-5 * np.maximum(np.abs(ref_b[l]), mag[l])
            assert np.all(np.abs(N(got[l]) - ref_b[l]) <= tol), (deterministic, l)
            if deterministic:   # bit-identical except the ONE pixel per image under the zero-padded rows (their taps
                same = (N(got[l]) == ref_b[l]).all(-1)      # are (0,0): pre-reduced, DESIGN.md "Backward")
                assert (~same).sum() <= B, (l, int((~same).sum()))
                assert same[:, 1:, :].all() and same[:, :, 1:].all(), l
        del got
        torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------------
def _stage(x, cfg, dev, B):
    from maskrcnn_tf2_b200.layers import DetectedBoxesExtraction, DetectionLayer, ProposalLayer, PyramidROIAlign
    fm = [T(f, dev) for f in x["feature_maps"]]
    meta = T(x["image_meta"], dev)
    rois = ProposalLayer(cfg["post_nms_rois_inference"], cfg)([T(x["rpn_probs"], dev), T(x["rpn_bbox"], dev),
                                                              T(x["anchors"], dev)])
    pooled = PyramidROIAlign([7, 7], name="roi_align_classifier")([rois, meta] + fm)
    det = DetectionLayer(cfg["post_nms_rois_inference"], cfg["detection_min_confidence"],
                         cfg["detection_max_instances"], cfg["detection_nms_threshold"], cfg["bbox_std_dev"], B, B)(
        [rois, T(x["mrcnn_class"], dev), T(x["mrcnn_bbox"], dev), meta])
    mask_pooled = PyramidROIAlign([14, 14], name="roi_align_mask")([DetectedBoxesExtraction(cfg)(det), meta] + fm)
    return dict(rois=rois, pooled=pooled, detections=det, mask_pooled=mask_pooled)


def _oracle_stage(orc, x, cfg, S):
    r = orc.proposal_layer(x["rpn_probs"], x["rpn_bbox"], x["anchors"], cfg["pre_nms_limit"],
                           cfg["post_nms_rois_inference"], SD, cfg["rpn_nms_threshold"])["proposals"]
    p7 = orc.pyramid_roi_align(r, float(S), float(S), x["feature_maps"], (7, 7))["out"]
    d = orc.detection_layer(r, x["mrcnn_class"], x["mrcnn_bbox"], x["image_meta"], SD, cfg["detection_min_confidence"],
                            cfg["detection_max_instances"], cfg["detection_nms_threshold"])["detections"]
    p14 = orc.pyramid_roi_align(np.ascontiguousarray(d[..., :4]), float(S), float(S), x["feature_maps"], (14, 14))["out"]
    return dict(rois=r, pooled=p7, detections=d, mask_pooled=p14)


@pytest.mark.parametrize("lo,n", [(0, 64), (0, 32), (32, 32), (0, 16), (16, 16), (0, 8), (8, 8), (56, 8)])
def test_config5_batch64_and_its_shards_first_appearance_mode(orc, dev, lo, n):
    """configs[4]: 64 COCO-shape images on one replica, and the shards of the 2 / 4 / 8-GPU split, every one in the
    reference's own map mode (Q2): each shard is compared with the oracle run on that shard, and (where a digest was
    generated) with the reference's own layer code run on that shard."""
    from maskrcnn_tf2_b200 import make_config, synth
    S = 1024
    x = synth.inference_batch(5, n, img_size=S, num_classes=81, regime="clustered", first_image=lo)
    cfg = make_config(img_size=S, num_classes=81, batch_size=n)
    got = _stage(x, cfg, dev, n)
    ref = _oracle_stage(orc, x, cfg, S)
    rec = digests()["config5"]["shards"].get(f"{lo}+{n}")
    for k in ("rois", "pooled", "detections", "mask_pooled"):
        g = N(got[k])
        assert np.array_equal(g, ref[k]), k
        if rec is not None:
            assert sha(g) == rec[k], k
    del got
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------------
def _threshold_pairs(thr, n_scales=20):
    """Pairs (A, B) sharing the corner at the origin, same height, widths a and b ~ a * thr: b is walked ulp by ulp until
    TF's fp32 IoU of the pair is exactly the fp32 threshold, or exactly 1 / 2 ulp above / below it (cycling through
    those five targets).  Scales shrink by 4x per pair and the four quadrants are used, so that members of different
    pairs overlap by at most IoU 1/16 / thr < 0.3."""
    t32 = np.float32(thr)
    up = lambda v, n: v if n == 0 else up(np.nextafter(v, np.float32(np.inf if n > 0 else -np.inf), dtype=np.float32),
                                          n - (1 if n > 0 else -1))
    targets = [up(t32, d) for d in (0, 1, -1, 2, -2)]
    pairs = []
    for s in range(n_scales):
        for q, (sy, sx) in enumerate(((1, 1), (1, -1), (-1, 1), (-1, -1))):
            scale = np.float32(2.0 ** (-2 * s))
            h, a = np.float32(0.25) * scale, np.float32(0.625) * scale
            want = targets[(4 * s + q) % len(targets)]
            A = np.array([0.0, 0.0, sy * h, sx * a], np.float32)
            best = None
            for k in sorted(range(-40, 41), key=abs):
                Bx = np.array([0.0, 0.0, sy * h, sx * up(np.float32(a * t32), k)], np.float32)
                got = _tf_iou_f32(A, Bx)[2]
                if got == want:
                    best = Bx
                    break
                if best is None:
                    best = Bx
            pairs.append((A, best))
    return pairs


def _tf_iou_f32(bi, bj):
    f = np.float32
    yi0, yi1 = min(bi[0], bi[2]), max(bi[0], bi[2]); xi0, xi1 = min(bi[1], bi[3]), max(bi[1], bi[3])
    yj0, yj1 = min(bj[0], bj[2]), max(bj[0], bj[2]); xj0, xj1 = min(bj[1], bj[3]), max(bj[1], bj[3])
    ai, aj = f(f(yi1 - yi0) * f(xi1 - xi0)), f(f(yj1 - yj0) * f(xj1 - xj0))
    ih = max(f(min(yi1, yj1) - max(yi0, yj0)), f(0)); iw = max(f(min(xi1, xj1) - max(xi0, xj0)), f(0))
    inter = f(ih * iw)
    uni = f(f(ai + aj) - inter)
    return inter, uni, f(inter / uni)


@pytest.mark.parametrize("thr", [0.3, 0.5, 0.7])
def test_nms_pairs_exactly_at_the_threshold_on_the_cluster_path(orc, dev, thr):
    from conftest import random_boxes
    from maskrcnn_tf2_b200 import functional as F
    rng = np.random.default_rng(int(thr * 100))
    M, B = 6000, 2
    pairs = _threshold_pairs(thr)
    n_band, n_equal = 0, 0
    for A, Bx in pairs:      # the test is only meaningful if these pairs land inside the kernel's screening band
        inter, uni, q = _tf_iou_f32(A, Bx)
        d = abs(float(inter) - float(np.float32(thr)) * float(uni))
        n_band += d <= float(uni) * thr * 2.0 ** -20
        n_equal += (q == np.float32(thr))
    assert n_band >= 60 and n_equal >= 10
    boxes = np.empty((B, M, 4), np.float32)
    scores = np.empty((B, M), np.float32)
    for b in range(B):
        fill = random_boxes(rng, M, clusters=60) * np.float32(4.0) + np.float32(4.0)     # far away from the origin
        order = rng.permutation(M)
        rank_score = (1.0 - np.arange(M) / M).astype(np.float32)                        # rank r -> score, descending
        boxes[b] = fill
        scores[b, order] = rank_score                                                    # box order[r] has rank r
        # A of pair p at rank ra, B at rank rb: same tile, next tile, 2..40 tiles apart (far path, other CTAs' share)
        gaps = [1, 7, 64, 65, 130, 640, 1280, 2560]
        for p, (A, Bx) in enumerate(pairs):
            ra = 40 * p + (3 if b == 0 else 11)
            rb = min(ra + gaps[p % len(gaps)], M - 1 - p)
            boxes[b, order[ra]] = A
            boxes[b, order[rb]] = Bx
    keep, count = F.nms(T(boxes, dev), T(scores, dev), 3000, thr)
    keep, count = N(keep), N(count)
    for b in range(B):
        ref = orc.nms(boxes[b], scores[b], 3000, thr)
        assert count[b] == len(ref)
        assert np.array_equal(keep[b, :len(ref)], ref)
    # same problem through ProposalLayer's entry (boxes arrive sorted; proposal epilogue)
    keep2, _ = F.nms(T(boxes, dev), T(scores, dev), 1000, thr)
    for b in range(B):
        assert np.array_equal(N(keep2)[b], keep[b, :1000])
