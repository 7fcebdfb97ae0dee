"""The oracle and the CUDA path against outputs of the REFERENCE'S OWN LAYER CODE.

tests/golden/reference_layers_golden.npz was written by tests/golden/make_reference_layers_golden.py, which loads
/root/reference/src/layers/mrcnn_layers.py and runs its ProposalLayer / PyramidROIAlign / DetectionLayer `call` methods
unmodified on a numpy stand-in for the `tf.*` functions they use (TensorFlow is not installable here).  That pins
everything the layers do with the ops -- arithmetic order, per-image slicing, the level formula (Q1), the
first-appearance map table and re-sort (Q2), the single class-agnostic NMS (Q3), intersections, padding, window
normalisation -- to the reference's own Python; the bodies of top_k / non_max_suppression / crop_and_resize in that
stand-in are an independent numpy restatement of the TF kernels (still not TensorFlow itself).  Bit-exact, all of it.
"""
import os

import numpy as np
import pytest
import torch

SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)


@pytest.fixture(scope="module")
def G():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_layers_golden.npz"))
    return {k: g[k] for k in g.files}


def _maps(G):
    return [G[f"fmap{i}"] for i in range(4)]


def test_golden_exercises_the_interesting_paths(G):
    kept400 = (G["rois_p400"] != 0).any(-1).sum(-1)
    assert np.all(kept400 < 400) and np.all(kept400 > 100)                  # zero padding of L:229-230 is exercised
    dets = (G["detections"][..., 4] > 0).sum(-1)
    assert dets[0] > 0 and dets[-1] == 0                                    # one image ends with no detection at all
    assert (G["detections_noconf"][..., 4] > 0).sum() > dets.sum()          # the confidence filter removed something
    assert len(np.unique(np.round(G["rpn_probs"][1, :, 1] * 64))) <= 65     # image 1: heavy score ties


def test_oracle_reproduces_the_reference_layers(orc, G):
    S, K, P, D = float(G["img_size"]), int(G["pre_nms_limit"]), int(G["proposal_count"]), int(G["max_instances"])
    r = orc.proposal_layer(G["rpn_probs"], G["rpn_bbox"], G["anchors"], K, P, SD, 0.7)
    assert np.array_equal(r["proposals"], G["rois"])
    r4 = orc.proposal_layer(G["rpn_probs"], G["rpn_bbox"], G["anchors"], K, 400, SD, 0.7)
    assert np.array_equal(r4["proposals"], G["rois_p400"])
    assert np.array_equal(orc.pyramid_roi_align(G["rois"], S, S, _maps(G), (7, 7))["out"], G["pooled"])
    d = orc.detection_layer(G["rois"], G["mrcnn_class"], G["mrcnn_bbox"], G["image_meta"], SD, 0.7, D, 0.3)
    assert np.array_equal(d["detections"], G["detections"])
    d0 = orc.detection_layer(G["rois"], G["mrcnn_class"], G["mrcnn_bbox"], G["image_meta"], SD, 0, D, 0.3)
    assert np.array_equal(d0["detections"], G["detections_noconf"])
    boxes = np.ascontiguousarray(G["detections"][..., :4])
    assert np.array_equal(orc.pyramid_roi_align(boxes, S, S, _maps(G), (14, 14))["out"], G["mask_pooled"])
    q = orc.pyramid_roi_align(G["q2_boxes"], S, S, _maps(G), (3, 5))
    assert np.array_equal(q["out"], G["q2_pooled"])
    assert q["roi_map"][0, 0] == 0 and q["roi_map"][0, 1] == 1              # Q2: first-seen level takes map 0


@pytest.mark.gpu
def test_cuda_layers_reproduce_the_reference_layers(G, dev):
    from maskrcnn_tf2_b200 import make_config
    from maskrcnn_tf2_b200.layers import DetectionLayer, ProposalLayer, PyramidROIAlign
    S, NC, B = int(G["img_size"]), int(G["num_classes"]), G["rois"].shape[0]
    K, P, D = int(G["pre_nms_limit"]), int(G["proposal_count"]), int(G["max_instances"])
    cfg = make_config(img_size=S, num_classes=NC, batch_size=B, pre_nms_limit=K)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    fm = [t(f) for f in _maps(G)]
    meta = t(G["image_meta"])
    inputs = [t(G["rpn_probs"]), t(G["rpn_bbox"]), t(G["anchors"])]
    rois = ProposalLayer(P, cfg)(inputs)
    assert np.array_equal(rois.cpu().numpy(), G["rois"])
    assert np.array_equal(ProposalLayer(400, cfg)(inputs).cpu().numpy(), G["rois_p400"])
    pooled = PyramidROIAlign([7, 7], name="roi_align_classifier")([rois, meta] + fm)
    assert np.array_equal(pooled.cpu().numpy(), G["pooled"])
    for conf, key in ((0.7, "detections"), (0, "detections_noconf")):
        det = DetectionLayer(P, conf, D, 0.3, cfg["bbox_std_dev"], B, B)([rois, t(G["mrcnn_class"]),
                                                                          t(G["mrcnn_bbox"]), meta])
        assert np.array_equal(det.cpu().numpy(), G[key]), key
    boxes = t(G["detections"][..., :4])
    mask_pooled = PyramidROIAlign([14, 14], name="roi_align_mask")([boxes, meta] + fm)
    assert np.array_equal(mask_pooled.cpu().numpy(), G["mask_pooled"])
    q2 = PyramidROIAlign([3, 5])([t(G["q2_boxes"]), meta] + fm)
    assert np.array_equal(q2.cpu().numpy(), G["q2_pooled"])


# ---- the training layer: DetectionTargetLayer.call -> detection_targets_graph, executed from the reference too -------
@pytest.fixture(scope="module")
def GT():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_target_layer_golden.npz"))
    return {k: g[k] for k in g.files}


def _target_args(GT, tag):
    return (GT["proposals"], GT["gt_class_ids"], GT["gt_boxes"], GT[f"gt_masks_{tag}"].astype(np.uint8), GT["rand_keys"],
            int(GT["train_rois_per_image"]), float(GT["roi_positive_ratio"]), SD, tuple(int(v) for v in GT["mask_shape"]))


@pytest.mark.parametrize("tag,mini", [("full", False), ("mini", True)])
def test_oracle_reproduces_the_reference_detection_target_layer(orc, GT, tag, mini):
    """tf.random.shuffle is unseeded in the reference (any permutation is its behaviour); the generator's stand-in draws
    the permutation the B200 layer derives from the injected keys, so the whole of L:844-967 is compared: zero trimming
    in the middle of the lists, the crowd box, `>= 0.5` / `< 0.5` / `< 0.001`, int(T*0.33) = 10 positives and
    int32(fp32(1/0.33)*10) - 10 = 20 negatives, first-max GT assignment, deltas / std, the 14x14 mask targets
    (full-size and mini-mask coordinates), padding."""
    r = orc.detection_target_layer(*_target_args(GT, tag), use_mini_masks=mini)
    for k in ("rois", "class_ids", "deltas", "masks"):
        assert np.array_equal(r[k], GT[f"{k}_{tag}"]), k
    assert np.array_equal(r["counts"], np.array([[10, 20]] * 3))
    assert set(np.unique(GT[f"masks_{tag}"])) == {0.0, 1.0}              # tf.round of the bilinear samples, L:954


@pytest.mark.gpu
@pytest.mark.parametrize("tag,mini", [("full", False), ("mini", True)])
def test_cuda_reproduces_the_reference_detection_target_layer(GT, dev, tag, mini):
    from maskrcnn_tf2_b200 import functional as F
    props, gtc, gtb, masks, keys, T_, ratio, sd, mshape = _target_args(GT, tag)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = F.detection_target_forward(t(props), t(gtc), t(gtb), t(masks), t(keys.view(np.int32)), T_, ratio, sd, mshape,
                                     use_mini_masks=mini)
    for got, k in zip(out, ("rois", "class_ids", "deltas", "masks")):
        assert np.array_equal(got.cpu().numpy(), GT[f"{k}_{tag}"]), k


# ---- the stand-in's own kernel restatements against the oracle's (two independent restatements of the TF kernels) ----
@pytest.fixture(scope="module")
def numpy_tf():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_layers_golden.py")
    spec = importlib.util.spec_from_file_location("make_reference_layers_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)            # defines the stand-in; touches /root/reference only when build() is called
    return mod.make_numpy_tf()


@pytest.mark.parametrize("seed", range(6))
def test_stand_in_kernels_agree_with_the_oracle_bit_for_bit(orc, numpy_tf, seed):
    from conftest import random_boxes
    tf = numpy_tf
    rng = np.random.default_rng(8800 + seed)
    m = int(rng.integers(5, 400))
    boxes = random_boxes(rng, m, clusters=int(rng.choice([0, 3])))
    boxes[1] = boxes[0]
    boxes[2, 2] = boxes[2, 0]                                             # zero area
    boxes[3] = boxes[3][[2, 3, 0, 1]]                                     # flipped corners
    scores = (np.round(rng.uniform(0, 1, m) * 32) / 32).astype(np.float32)
    scores[4] = -np.inf
    for thr in (0.0, 0.3, 0.7, 1.0):
        mo = int(rng.integers(1, m + 5))
        assert np.array_equal(np.asarray(tf.image.non_max_suppression(boxes, scores, mo, thr)),
                              orc.nms(boxes, scores, mo, thr)), (m, mo, thr)
    k = int(rng.integers(1, m + 1))
    assert np.array_equal(np.asarray(tf.nn.top_k(scores, k).indices), orc.topk(scores, k))
    H, W, C, n = int(rng.integers(1, 20)), int(rng.integers(1, 20)), 3, 9
    img = rng.standard_normal((2, H, W, C)).astype(np.float32)
    cb = rng.uniform(-0.3, 1.3, (n, 4)).astype(np.float32)
    cb[0] = [0, 0, 1, 1]
    cb[1] = 0
    ind = rng.integers(0, 2, n).astype(np.int32)
    for crop in ((7, 7), (1, 1), (2, 5), (14, 3)):
        assert np.array_equal(np.asarray(tf.image.crop_and_resize(img, cb, ind, crop)),
                              orc.crop_and_resize(img, cb, ind, crop)), (H, W, crop)


# ---- BASELINE.json's full sizes: SHA-256 of the reference layers' outputs at COCO shape ---------------------------
def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def full_size():
    import json
    from maskrcnn_tf2_b200 import synth
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_layers_full_size_sha256.json")
    rec = json.load(open(path))
    x = synth.inference_batch(2, 2, img_size=1024, num_classes=81, regime="clustered", n_rois=1000, channels=256)
    for k, want in rec["input_sha256"].items():          # the inputs are regenerated from seeds, not stored
        assert _sha(x[k]) == want, f"synthetic input {k} is not the array the digests were made from"
    assert [_sha(f) for f in x["feature_maps"]] == rec["fmap_sha256"]
    return rec, x


def test_oracle_reproduces_the_reference_layers_at_coco_shape(orc, full_size):
    """1024^2, A = 261 888 -> top-6000 -> NMS 0.7 -> 1000 proposals -> ROIAlign 7x7 over 256-channel maps -> 81-class
    DetectionLayer -> ROIAlign 14x14: digests of the reference's own layer code (make_reference_layers_golden.py)."""
    rec, x = full_size
    r = orc.proposal_layer(x["rpn_probs"], x["rpn_bbox"], x["anchors"], 6000, 1000, SD, 0.7)
    assert _sha(r["proposals"]) == rec["sha256"]["rois"]
    pooled = orc.pyramid_roi_align(r["proposals"], 1024.0, 1024.0, x["feature_maps"], (7, 7))["out"]
    assert _sha(pooled) == rec["sha256"]["pooled"]
    det = orc.detection_layer(r["proposals"], x["mrcnn_class"], x["mrcnn_bbox"], x["image_meta"], SD, 0.7, 100, 0.3)
    assert _sha(det["detections"]) == rec["sha256"]["detections"] and list(det["count"]) == rec["detections"]
    mask = orc.pyramid_roi_align(np.ascontiguousarray(det["detections"][..., :4]), 1024.0, 1024.0, x["feature_maps"],
                                 (14, 14))["out"]
    assert _sha(mask) == rec["sha256"]["mask_pooled"]


@pytest.mark.gpu
def test_cuda_layers_reproduce_the_reference_layers_at_coco_shape(full_size, dev):
    from maskrcnn_tf2_b200 import make_config
    from maskrcnn_tf2_b200.layers import DetectionLayer, ProposalLayer, PyramidROIAlign
    rec, x = full_size
    B = 2
    cfg = make_config(img_size=1024, num_classes=81, batch_size=B)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    fm = [t(f) for f in x["feature_maps"]]
    meta = t(x["image_meta"])
    rois = ProposalLayer(1000, cfg)([t(x["rpn_probs"]), t(x["rpn_bbox"]), t(x["anchors"])])
    assert _sha(rois.cpu().numpy()) == rec["sha256"]["rois"]
    pooled = PyramidROIAlign([7, 7], name="roi_align_classifier")([rois, meta] + fm)
    assert _sha(pooled.cpu().numpy()) == rec["sha256"]["pooled"]
    det = DetectionLayer(1000, 0.7, 100, 0.3, cfg["bbox_std_dev"], B, B)([rois, t(x["mrcnn_class"]), t(x["mrcnn_bbox"]),
                                                                          meta])
    assert _sha(det.cpu().numpy()) == rec["sha256"]["detections"]
    mask = PyramidROIAlign([14, 14], name="roi_align_mask")([det[..., :4].contiguous(), meta] + fm)
    assert _sha(mask.cpu().numpy()) == rec["sha256"]["mask_pooled"]


def _training_case():
    import json
    from maskrcnn_tf2_b200 import synth
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_layers_full_size_sha256.json")
    rec = json.load(open(path))["training"]
    x = synth.inference_batch(3, 2, img_size=512, regime="clustered")
    g = synth.training_targets_batch(3, 2, img_size=512)
    keys = np.random.default_rng(7).integers(0, 2 ** 32, (2, 2000), dtype=np.uint64).astype(np.uint32)
    have = {"rpn_probs": x["rpn_probs"], "rpn_bbox": x["rpn_bbox"], "gt_class_ids": g["gt_class_ids"],
            "gt_boxes": g["gt_boxes"], "gt_masks": g["gt_masks"], "keys": keys}
    for k, want in rec["input_sha256"].items():
        assert _sha(have[k]) == want, f"synthetic input {k} is not the array the digests were made from"
    return rec, x, g, keys


def test_oracle_reproduces_the_reference_training_step_at_config3_shape(orc):
    """ProposalLayer(2000) -> DetectionTargetLayer(T=200, full-size 512^2 masks, 100 GT slots) of the reference's own code
    on the inputs of tests/test_gpu_configs.py::test_config2_training_step_*: that GPU test holds the CUDA path equal
    to the oracle on exactly these arrays, so CUDA == oracle == reference code."""
    rec, x, g, keys = _training_case()
    props = orc.proposal_layer(x["rpn_probs"], x["rpn_bbox"], x["anchors"], 6000, 2000, SD, 0.7)["proposals"]
    assert _sha(props) == rec["sha256"]["proposals"]
    r = orc.detection_target_layer(props, g["gt_class_ids"], g["gt_boxes"], g["gt_masks"], keys, 200, 0.33, SD, (28, 28))
    for k in ("rois", "class_ids", "deltas", "masks"):
        assert _sha(r[k]) == rec["sha256"][k], k
    assert [int(c) for c in r["counts"][:, 0]] == rec["positives"] == [66, 66]


@pytest.mark.parametrize("mini", [False, True])
def test_oracle_reproduces_the_reference_training_step_at_full_config3_size(orc, mini):
    """BASELINE.json configs[2] at FULL size (1024^2, 8 images, 2000 proposals, T = 200, full 1024^2 masks / 32x32
    mini-masks) + PyramidROIAlign 7x7 and 14x14 on the target ROIs: digests of the reference's own layer code.  The GPU
    twin is tests/test_gpu_full_size.py::test_config3_full_size_training_step."""
    import json
    from maskrcnn_tf2_b200 import synth
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_layers_full_size_sha256.json")
    rec = json.load(open(path))["training_full_mini" if mini else "training_full"]
    B = 8
    x = synth.inference_batch(3, B, img_size=1024, regime="clustered")
    g = synth.training_targets_batch(3, B, img_size=1024, mini_mask=(32, 32) if mini else None)
    keys = np.random.default_rng(9).integers(0, 2 ** 32, (B, 2000), dtype=np.uint64).astype(np.uint32)
    have = {"rpn_probs": x["rpn_probs"], "rpn_bbox": x["rpn_bbox"], "gt_class_ids": g["gt_class_ids"],
            "gt_boxes": g["gt_boxes"], "gt_masks": g["gt_masks"], "keys": keys}
    for k, want in rec["input_sha256"].items():
        assert _sha(have[k]) == want, k
    props = orc.proposal_layer(x["rpn_probs"], x["rpn_bbox"], x["anchors"], 6000, 2000, SD, 0.7)["proposals"]
    assert _sha(props) == rec["sha256"]["proposals"]
    r = orc.detection_target_layer(props, g["gt_class_ids"], g["gt_boxes"], g["gt_masks"], keys, 200, 0.33, SD, (28, 28),
                                   use_mini_masks=mini)
    for k in ("rois", "class_ids", "deltas", "masks"):
        assert _sha(r[k]) == rec["sha256"][k], k
    if not mini:
        for pool, name in (((7, 7), "pooled7"), ((14, 14), "pooled14")):
            out = orc.pyramid_roi_align(r["rois"], 1024.0, 1024.0, x["feature_maps"], pool)["out"]
            assert _sha(out) == rec["sha256"][name], name


def test_oracle_reproduces_the_reference_layers_on_a_config5_shard(orc):
    """BASELINE.json configs[4]: the 8-image shard an 8-GPU split of the batch of 64 gives rank 1 (images 8..15), in the
    reference's first-appearance map mode; digests of the reference's own layer code.  All shard sizes run on the GPU
    (tests/test_gpu_full_size.py::test_config5_*)."""
    import json
    from maskrcnn_tf2_b200 import synth
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_layers_full_size_sha256.json")
    rec = json.load(open(path))["config5"]["shards"]["8+8"]
    x = synth.inference_batch(5, 8, img_size=1024, num_classes=81, regime="clustered", first_image=8)
    r = orc.proposal_layer(x["rpn_probs"], x["rpn_bbox"], x["anchors"], 6000, 1000, SD, 0.7)["proposals"]
    assert _sha(r) == rec["rois"]
    assert _sha(orc.pyramid_roi_align(r, 1024.0, 1024.0, x["feature_maps"], (7, 7))["out"]) == rec["pooled"]
    d = orc.detection_layer(r, x["mrcnn_class"], x["mrcnn_bbox"], x["image_meta"], SD, 0.7, 100, 0.3)["detections"]
    assert _sha(d) == rec["detections"]
    m = orc.pyramid_roi_align(np.ascontiguousarray(d[..., :4]), 1024.0, 1024.0, x["feature_maps"], (14, 14))["out"]
    assert _sha(m) == rec["mask_pooled"]
