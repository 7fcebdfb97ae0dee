"""Seeded random sweeps over ragged shapes: every launcher against the CPU oracle on shapes nobody picked by hand
(odd batch sizes, M / A / N that are not multiples of any tile, max_out above and below the survivor count, valid
counts from 0 to M, non-square feature maps, odd pool shapes, channel counts that only satisfy C % 4 == 0).  Same
bars as tests/test_gpu_parity.py: indices bit-exact, fp32 values bit-exact (inside 1e-5 rel / 1e-6 abs), atomically
accumulated gradients within tolerance.  Each case is small so the whole file runs in seconds."""
import os

import numpy as np
import pytest
import torch

from conftest import random_boxes

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
MULT = int(os.environ.get("MRCNN_FUZZ_MULT", "1"))     # MRCNN_FUZZ_MULT=10: the long sweep (profiles/r1_fuzz.md)


@pytest.fixture(scope="module")
def F():
    from maskrcnn_tf2_b200 import functional
    return functional


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


def _scores(rng, shape):
    kind = rng.integers(0, 4)
    if kind == 0:
        return rng.standard_normal(shape).astype(np.float32)
    if kind == 1:       # heavy ties
        return (np.round(rng.uniform(0, 1, shape) * 32) / 32).astype(np.float32)
    if kind == 2:       # probabilities saturating at both ends
        return (1 / (1 + np.exp(-rng.standard_normal(shape) * 20))).astype(np.float32)
    s = rng.uniform(0, 1, shape).astype(np.float32)
    s[rng.uniform(0, 1, shape) < 0.02] = -np.inf
    return s


@pytest.mark.parametrize("seed", range(24 * MULT))
def test_fuzz_topk(F, orc, dev, seed):
    rng = np.random.default_rng(7000 + seed)
    B = int(rng.integers(1, 6))
    A = int(rng.choice([rng.integers(1, 100), rng.integers(100, 9000), rng.integers(9000, 120000)]))
    K = int(rng.integers(1, min(A, 8192) + 1))
    s = _scores(rng, (B, A))
    idx = N(F.topk(T(s, dev), K))
    for b in range(B):
        assert np.array_equal(idx[b], orc.topk(s[b], K)), (B, A, K)


@pytest.mark.parametrize("seed", range(32 * MULT))
def test_fuzz_nms(F, orc, dev, seed):
    rng = np.random.default_rng(7100 + seed)
    B = int(rng.integers(1, 10))
    M = int(rng.choice([rng.integers(1, 130), rng.integers(130, 2100), rng.integers(2100, 8193)]))
    max_out = int(rng.choice([1, rng.integers(1, M + 1), M, M + 7, min(4 * M, 8192)]))
    thr = float(rng.choice([0.0, 0.3, 0.5, 0.7, 1.0, rng.uniform(0, 1)]))
    clusters = int(rng.choice([0, 1, 3, 40]))
    boxes = np.stack([random_boxes(rng, M, clusters=clusters) for _ in range(B)])
    if M > 8 and seed % 3 == 0:    # duplicates, zero-area and flipped boxes
        boxes[:, M // 2:M // 2 + M // 8] = boxes[:, :M // 8]
        boxes[:, 1::7, 2] = boxes[:, 1::7, 0]
        boxes[:, 3::11] = boxes[:, 3::11][:, :, [2, 3, 0, 1]]
    scores = _scores(rng, (B, M))
    valid = None
    if seed % 2:
        valid = np.minimum(rng.integers(0, M + 1, B), M).astype(np.int32)
        valid[rng.integers(0, B)] = rng.choice([0, M])
    keep, count = F.nms(T(boxes, dev), T(scores, dev), max_out, thr, None if valid is None else T(valid, dev))
    keep, count = N(keep), N(count)
    for b in range(B):
        n = M if valid is None else int(valid[b])
        ref = orc.nms(boxes[b, :n], scores[b, :n], max_out, thr)
        assert count[b] == len(ref), (B, M, max_out, thr, n)
        assert np.array_equal(keep[b, :len(ref)], ref), (B, M, max_out, thr, n)
        assert np.all(keep[b, len(ref):] == -1)


def _roi_boxes(rng, B, Nr, img):
    side = np.exp(rng.uniform(np.log(4), np.log(img), (B, Nr))) / img
    ar = np.exp(rng.uniform(-1.0, 1.0, (B, Nr)))
    h, w = np.minimum(side * ar, 1.0), np.minimum(side / ar, 1.0)
    y1, x1 = rng.uniform(0, 1, (B, Nr)) * (1 - h), rng.uniform(0, 1, (B, Nr)) * (1 - w)
    boxes = np.stack([y1, x1, y1 + h, x1 + w], -1).astype(np.float32)
    wild = int(rng.integers(0, Nr // 3 + 1))
    if wild:
        boxes[:, :wild] += rng.uniform(-0.6, 0.6, (B, wild, 4)).astype(np.float32)
    pad = int(rng.integers(0, Nr // 2 + 1))
    if pad:
        boxes[:, Nr - pad:] = 0.0
    return boxes


@pytest.mark.parametrize("seed", range(24 * MULT))
def test_fuzz_roialign_forward_and_backward(F, orc, dev, seed):
    from maskrcnn_tf2_b200 import synth
    rng = np.random.default_rng(7200 + seed)
    B, Nr = int(rng.integers(1, 5)), int(rng.integers(1, 90))
    C = int(rng.choice([4, 12, 64, 100, 256, 260]))
    pool = (int(rng.choice([1, 2, 7, 14])), int(rng.choice([1, 3, 7, 14])))
    img = float(rng.choice([256, 512, 1024]))
    base = int(rng.choice([8, 24, 40]))
    hw = [(max(base >> l, 1) + int(rng.integers(0, 3)), max(base >> l, 1) + int(rng.integers(0, 3))) for l in range(4)]
    mode = int(rng.integers(0, 2))
    boxes = _roi_boxes(rng, B, Nr, img)
    fm = [rng.standard_normal((B, h, w, C)).astype(np.float32) for h, w in hw]
    meta = synth.image_meta(B, int(img), 81)
    ref = orc.pyramid_roi_align(boxes, img, img, fm, pool, map_mode=mode)
    out, roi_map = F.roialign_forward(T(boxes, dev), T(meta, dev), [T(f, dev) for f in fm], pool, map_mode=mode)
    assert np.array_equal(N(roi_map), ref["roi_map"]), (B, Nr, C, pool, hw, mode)
    assert np.array_equal(N(out), ref["out"]), (B, Nr, C, pool, hw, mode)
    # gradient: sequential-order (deterministic) and atomic variants against the oracle's sequential sum
    g = rng.standard_normal((B, Nr) + pool + (C,)).astype(np.float32)
    shapes = [f.shape for f in fm]
    gref = orc.pyramid_roi_align_grad(g, boxes, img, img, shapes, map_mode=mode)
    mag = orc.pyramid_roi_align_grad(np.abs(g), boxes, img, img, shapes, map_mode=mode)
    for deterministic in (True, False):
        grads = F.roialign_backward(T(g, dev), T(boxes, dev), roi_map, shapes, deterministic=deterministic)
        for l in range(4):
            # tolerance relative to the accumulated magnitude: the summation order differs from the oracle's
            err = np.abs(N(grads[l]) - gref[l])
            assert np.all(err <= ATOL + RTOL * mag[l]), (deterministic, l, float(err.max()), (B, Nr, C, pool, hw))


@pytest.mark.parametrize("seed", range(16 * MULT))
def test_fuzz_detection_layer(F, orc, dev, seed):
    from maskrcnn_tf2_b200 import synth
    rng = np.random.default_rng(7300 + seed)
    B, Nr, NC = int(rng.integers(1, 7)), int(rng.integers(1, 1500)), int(rng.choice([2, 3, 81, 100]))
    D = int(rng.choice([1, 10, 100, 300]))
    min_conf = float(rng.choice([0.0, 0.3, 0.7, 0.99]))
    thr = float(rng.choice([0.0, 0.3, 0.5, 1.0]))
    rois = np.stack([random_boxes(rng, Nr, clusters=int(rng.choice([0, 4]))) for _ in range(B)])
    z = float(rng.choice([1.0, 3.0, 8.0])) * rng.standard_normal((B, Nr, NC))
    probs = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(np.float32)
    if seed % 4 == 0:
        probs[0] = 0.0
        probs[0, :, 0] = 1.0                      # one image with background only
    deltas = rng.standard_normal((B, Nr, NC, 4)).astype(np.float32)
    meta = synth.image_meta(B, 1024, NC)
    if seed % 3 == 0:                             # a window smaller than the image (padded input)
        meta[:, 7:11] = np.array([100, 50, 900, 1000], np.float32)
    ref = orc.detection_layer(rois, probs, deltas, meta, SD, min_conf, D, thr)
    got = F.detection_forward(T(rois, dev), T(probs, dev), T(deltas, dev), T(meta, dev), SD, min_conf, D, thr)
    det = got[0] if isinstance(got, (tuple, list)) else got
    assert np.array_equal(N(det), ref["detections"]), (B, Nr, NC, D, min_conf, thr)


@pytest.mark.parametrize("seed", range(12 * MULT))
def test_fuzz_proposal_layer(F, orc, dev, seed):
    rng = np.random.default_rng(7400 + seed)
    B = int(rng.integers(1, 5))
    A = int(rng.choice([rng.integers(1, 400), rng.integers(400, 20000)]))
    K = int(rng.choice([6000, rng.integers(1, 8192)]))
    P = int(rng.choice([1, 50, 1000, 2000]))
    thr = float(rng.choice([0.5, 0.7, 0.9]))
    anchors1 = random_boxes(rng, A, clusters=int(rng.choice([0, 8])))
    anchors = np.ascontiguousarray(np.broadcast_to(anchors1, (B, A, 4)))
    fg = _scores(rng, (B, A))
    fg = np.where(np.isfinite(fg), fg, 0.0).astype(np.float32)
    probs = np.stack([1 - fg, fg], -1).astype(np.float32)
    bbox = (rng.standard_normal((B, A, 4)) * float(rng.choice([0.1, 1.0, 4.0]))).astype(np.float32)
    ref = orc.proposal_layer(probs, bbox, anchors, K, P, SD, thr)
    got = F.proposal_forward(T(probs, dev), T(bbox, dev), T(anchors, dev), K, P, SD, thr)
    prop = got[0] if isinstance(got, (tuple, list)) else got["proposals"] if isinstance(got, dict) else got
    assert np.array_equal(N(prop), ref["proposals"]), (B, A, K, P, thr)


@pytest.mark.parametrize("seed", range(12 * MULT))
def test_fuzz_detection_target_layer(F, orc, dev, seed):
    rng = np.random.default_rng(7500 + seed)
    B, P, G = int(rng.integers(1, 5)), int(rng.integers(8, 3000)), int(rng.choice([1, 7, 100]))
    T_ = int(rng.choice([1, 17, 200, 512]))
    mini = bool(seed % 2)
    MH = 28 if mini else int(rng.choice([20, 64, 130]))
    ratio = float(rng.choice([0.33, 0.5, 0.25]))
    n_real = int(rng.integers(0, min(G, 12) + 1))
    props = np.stack([random_boxes(rng, P, min_size=0.04, max_size=0.5, clusters=int(rng.choice([0, 6])))
                      for _ in range(B)])
    props[:, P - P // 5:] = 0.0
    gtb = np.zeros((B, G, 4), np.float32)
    gtc = np.zeros((B, G), np.int32)
    for b in range(B):
        pick = rng.integers(0, max(P - P // 5, 1), n_real)
        gtb[b, :n_real] = props[b, pick] + rng.normal(0, 0.006, (n_real, 4)).astype(np.float32)
        gtc[b, :n_real] = rng.integers(1, 81, n_real) * rng.choice([1, 1, 1, -1], n_real)    # some crowds
    masks = (rng.uniform(0, 1, (B, MH, MH, G)) < 0.5).astype(np.uint8)
    keys = rng.integers(0, 2 ** 32, (B, P), dtype=np.uint64).astype(np.uint32)
    if seed % 3 == 0:
        keys[:, ::2] = keys[:, 1::2][:, :keys[:, ::2].shape[1]] if P % 2 == 0 else keys[:, ::2]   # tied keys
    ref = orc.detection_target_layer(props, gtc, gtb, masks, keys, T_, ratio, SD, (28, 28), use_mini_masks=mini)
    out = F.detection_target_forward(T(props, dev), T(gtc, dev), T(gtb, dev), T(masks, dev),
                                     T(keys.view(np.int32), dev), T_, ratio, SD, (28, 28), use_mini_masks=mini,
                                     return_counts=True)
    for got, name in zip(out, ["rois", "class_ids", "deltas", "masks", "counts"]):
        assert np.array_equal(N(got), ref[name], equal_nan=True), (name, B, P, G, T_, MH, mini, ratio, n_real)
