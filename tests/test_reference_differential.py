"""Differential test against the reference's own layer code, run live where /root/reference exists (this container;
skipped on the GPU box, which has no copy of the reference): random small cases through ProposalLayer /
PyramidROIAlign / DetectionLayer / DetectionTargetLayer of /root/reference/src/layers/mrcnn_layers.py -- executed
unmodified on the numpy stand-in for tf.* of tests/golden/make_reference_layers_golden.py -- and through the oracle,
compared bit for bit.  The committed fixtures of tests/test_reference_layers.py are one such case frozen; this sweeps
shapes, regimes, thresholds, windows, crowds, padding and mini-masks."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import random_boxes

pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/src/layers/mrcnn_layers.py"),
                                reason="needs the reference checkout (not present on the GPU box)")
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)


@pytest.fixture(scope="module")
def ref():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_layers_golden.py")
    spec = importlib.util.spec_from_file_location("make_reference_layers_golden", path)
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen, gen.load_reference_layers()


def _t(gen, a):
    return np.asarray(a).view(gen.T)


@pytest.mark.parametrize("seed", range(8))
def test_inference_layers_live_against_the_reference(orc, ref, seed):
    from maskrcnn_tf2_b200 import synth
    gen, L = ref
    rng = np.random.default_rng(9100 + seed)
    S = int(rng.choice([64, 128]))
    B, NC, C = int(rng.integers(1, 4)), int(rng.choice([2, 7])), 4
    K, P, D = int(rng.choice([60, 300, 6000])), int(rng.choice([10, 80])), int(rng.choice([5, 30]))
    thr = float(rng.choice([0.5, 0.7]))
    anchors1 = synth.pyramid_anchors(S)
    A = anchors1.shape[0]
    regime = str(rng.choice(["clustered", "iid", "sparse"]))
    probs, bbox = zip(*[synth.rpn_outputs(np.random.default_rng(seed * 10 + b), anchors1, regime, S) for b in range(B)])
    probs, bbox = np.stack(probs).astype(np.float32), np.stack(bbox).astype(np.float32)
    if seed % 2:
        probs[..., 1] = np.round(probs[..., 1] * 32) / 32            # ties
        probs[..., 0] = 1 - probs[..., 1]
    anchors = np.ascontiguousarray(np.broadcast_to(anchors1, (B, A, 4))).astype(np.float32)
    cfg = {"rpn_nms_threshold": thr, "pre_nms_limit": K, "images_per_gpu": B, "rpn_bbox_std_dev": SD, "bbox_std_dev": SD}
    t = lambda a: _t(gen, a)
    rois = np.asarray(L.ProposalLayer(proposal_count=P, config=cfg)([t(probs), t(bbox), t(anchors)]))
    assert np.array_equal(orc.proposal_layer(probs, bbox, anchors, K, P, SD, thr)["proposals"], rois)

    meta = synth.image_meta(B, S, NC).astype(np.float32)
    if seed % 3 == 0:
        meta[:, 7:11] = [3, 5, S - 9, S - 2]
    if seed % 2 and B > 1:                                  # Q6: only image 0's shape is used (L:514-515, L:600)
        meta[1:, 4:6] = [S // 2, 2 * S]
    fm = [rng.standard_normal((B, max(S // s, 1), max(S // s, 1), C)).astype(np.float32) for s in (4, 8, 16, 32)]
    boxes = rois.copy()
    boxes[:, ::5] = np.stack([random_boxes(rng, boxes[:, ::5].shape[1], min_size=0.02, max_size=0.9) for _ in range(B)])
    boxes[:, 1::7] += rng.uniform(-0.4, 0.4, boxes[:, 1::7].shape).astype(np.float32)     # partly outside the image
    for pool in ((7, 7), (int(rng.integers(1, 5)), int(rng.integers(1, 5)))):
        want = np.asarray(L.PyramidROIAlign(list(pool))([t(boxes), t(meta)] + [t(f) for f in fm]))
        assert np.array_equal(orc.pyramid_roi_align(boxes, float(S), float(S), fm, pool)["out"], want), pool

    z = float(rng.choice([2.0, 5.0])) * rng.standard_normal((B, P, NC))
    mc = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(np.float32)
    mb = rng.standard_normal((B, P, NC, 4)).astype(np.float32)
    conf = float(rng.choice([0.0, 0.5, 0.7]))
    dthr = float(rng.choice([0.3, 0.5]))
    det = np.asarray(L.DetectionLayer(proposals=P, detection_min_confidence=conf, detection_max_instances=D,
                                      detection_nms_threshold=dthr, bbox_std_dev=SD, images_per_gpu=B, batch_size=B)(
        [t(rois), t(mc), t(mb), t(meta)]))
    assert np.array_equal(orc.detection_layer(rois, mc, mb, meta, SD, conf, D, dthr)["detections"], det)


@pytest.mark.parametrize("seed", range(8))
def test_detection_target_layer_live_against_the_reference(orc, ref, seed):
    gen, L = ref
    rng = np.random.default_rng(9200 + seed)
    B, P, G = int(rng.integers(1, 4)), int(rng.integers(20, 250)), int(rng.choice([3, 10]))
    T_ = int(rng.choice([8, 33, 64]))
    ratio = float(rng.choice([0.33, 0.25, 0.5]))
    mini = bool(seed % 2)
    mh = 12 if mini else int(rng.choice([9, 24]))
    mshape = (int(rng.choice([7, 14])), int(rng.choice([7, 14])))
    props = np.stack([random_boxes(rng, P, min_size=0.05, max_size=0.5, clusters=int(rng.choice([0, 4])))
                      for _ in range(B)])
    props[:, P - P // 6:] = 0
    gtb = np.zeros((B, G, 4), np.float32)
    gtc = np.zeros((B, G), np.int32)
    for b in range(B):
        n_real = int(rng.integers(1, G + 1))
        rows = rng.choice(G, n_real, replace=False)                     # real rows scattered between zero rows
        pick = rng.integers(0, P - P // 6, n_real)
        gtb[b, rows] = np.clip(props[b, pick] + rng.normal(0, 0.01, (n_real, 4)).astype(np.float32), 0, 1)
        gtc[b, rows] = rng.integers(1, 81, n_real)
        if n_real > 1 and rng.uniform() < 0.5:
            gtc[b, rows[0]] *= -1                                       # crowd
    assert np.all((gtb[..., 2] > gtb[..., 0]) | (gtc == 0)) and np.all((gtb[..., 3] > gtb[..., 1]) | (gtc == 0))
    masks = rng.uniform(0, 1, (B, mh, mh, G)) < 0.5
    keys = rng.integers(0, 2 ** 32, (B, P), dtype=np.uint64).astype(np.uint32)
    cfg = {"train_rois_per_image": T_, "roi_positive_ratio": ratio, "use_mini_masks": mini, "mask_shape": mshape,
           "bbox_std_dev": SD, "images_per_gpu": B}
    gen.SHUFFLE["calls"] = 0
    gen.SHUFFLE["per_image"] = [(keys[b], np.flatnonzero(np.abs(props[b]).sum(1) != 0)) for b in range(B)]
    t = lambda a: _t(gen, a)
    rois, cls, deltas, tm = L.DetectionTargetLayer(cfg)([t(props), t(gtc), t(gtb), t(masks)])
    r = orc.detection_target_layer(props, gtc, gtb, masks.astype(np.uint8), keys, T_, ratio, SD, mshape,
                                   use_mini_masks=mini)
    assert np.array_equal(r["rois"], np.asarray(rois)) and np.array_equal(r["class_ids"], np.asarray(cls))
    assert np.array_equal(r["deltas"], np.asarray(deltas), equal_nan=True)
    assert np.array_equal(r["masks"], np.asarray(tm))


@pytest.mark.parametrize("S", [64, 128, 256])
def test_anchors_layer_live_against_the_reference(ref, S):
    """AnchorsLayer (L:104-145): numpy float64 pyramid -> Keras autocast to float32 -> NormBoxesLayer, broadcast to the
    batch.  The host-side synth.pyramid_anchors (what the benchmarks feed) must be the same bits; the device producer
    is compared with the same vectors in tests/test_reference_pins.py."""
    from maskrcnn_tf2_b200 import synth
    gen, L = ref
    cfg = {"image_shape": (S, S, 3), "img_size": S, "batch_size": 2, "backbone_strides": [4, 8, 16, 32, 64],
           "rpn_anchor_scales": (32, 64, 128, 256, 512), "rpn_anchor_ratios": [0.5, 1, 2], "rpn_anchor_stride": 1,
           "backbone": "resnet50"}
    got = np.asarray(L.AnchorsLayer(cfg, training=False).call())
    want = synth.pyramid_anchors(S)
    assert got.dtype == np.float32 and got.shape == (2,) + want.shape
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)


def test_full_size_digests_are_what_the_reference_layers_produce_now(ref):
    """Regenerates the COCO-shape run of the reference's layers live and compares with the committed digests (guards the
    committed JSON against drift of the generator or of numpy)."""
    import json
    gen, L = ref
    rec = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                      "reference_layers_full_size_sha256.json")))
    out = gen.run_full_size(L, gen.full_size_inputs())
    assert {k: gen.digest(v.astype(np.float32)) for k, v in out.items()} == rec["sha256"]
