"""Property tests of the CPU oracle (hypothesis; CPU only).  The reference holds no vectors for the TensorFlow
kernels' arithmetic ("parity unpinned"), so beside the hand-derived known answers (test_oracle_kat.py)
and the independent implementations (test_oracle_cross.py) the oracle must satisfy the defining properties of each
op on arbitrary inputs: greedy-NMS maximality, the top-k order relation, bilinearity and the adjoint identity of
crop_and_resize / its gradient, padding and ordering rules of the layers."""
import numpy as np
from hypothesis import given, settings, strategies as st

from conftest import random_boxes

SET = dict(max_examples=25, deadline=None, derandomize=True)
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)


def _rank_order(scores):
    idx = np.arange(scores.size)
    return np.lexsort((idx, -scores.astype(np.float64)))       # score desc, index asc


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), m=st.integers(1, 300), thr=st.sampled_from([0.0, 0.3, 0.5, 0.7, 1.0]),
       quant=st.booleans(), clusters=st.sampled_from([0, 1, 5]))
def test_nms_is_the_greedy_maximal_selection(orc, seed, m, thr, quant, clusters):
    rng = np.random.default_rng(seed)
    boxes = random_boxes(rng, m, clusters=clusters)
    if m > 6:
        boxes[1] = boxes[0]                                      # duplicate
        boxes[2, 2] = boxes[2, 0]                                # zero area
        boxes[3] = boxes[3][[2, 3, 0, 1]]                        # flipped corners
    scores = rng.uniform(0, 1, m).astype(np.float32)
    if quant:
        scores = (np.round(scores * 8) / 8).astype(np.float32)
    max_out = int(rng.integers(1, m + 3))
    keep = orc.nms(boxes, scores, max_out, thr)
    order = _rank_order(scores)
    rank = np.empty(m, int)
    rank[order] = np.arange(m)
    assert len(set(keep.tolist())) == len(keep) <= min(max_out, m)
    assert np.all(np.diff(rank[keep]) > 0)                       # selected in candidate order
    kept = set(keep.tolist())
    for a in range(len(keep)):                                   # no kept pair overlaps above the threshold
        for b in range(a):
            assert not orc.tf_iou(boxes, int(keep[a]), int(keep[b])) > thr
    last_rank = rank[keep[-1]] if len(keep) == max_out else m    # the sweep stops once max_out boxes are kept
    for i in range(m):                                           # maximality: every skipped candidate was suppressed
        if i in kept or rank[i] > last_rank:
            continue
        assert any(rank[k] < rank[i] and orc.tf_iou(boxes, i, int(k)) > thr for k in keep), i


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 5000), levels=st.sampled_from([0, 4, 64]))
def test_topk_is_the_prefix_of_the_total_order(orc, seed, n, levels):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal(n).astype(np.float32)
    if levels:
        s = (np.round(s * levels) / levels).astype(np.float32)   # ties (and both signed zeros)
    if n > 4:
        s[rng.integers(0, n)] = np.inf
        s[rng.integers(0, n)] = -np.inf
        s[rng.integers(0, n)] = -0.0
    k = int(rng.integers(1, n + 1))
    got = orc.topk(s, k)
    assert np.array_equal(got, _rank_order(s)[:k])               # -0.0 == 0.0 compare equal: index decides


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), h=st.integers(1, 12), w=st.integers(1, 12), ph=st.integers(1, 6),
       pw=st.integers(1, 6))
def test_crop_and_resize_is_linear_and_its_gradient_is_the_adjoint(orc, seed, h, w, ph, pw):
    rng = np.random.default_rng(seed)
    B, C, n = 2, 3, 7
    x, y = rng.standard_normal((2, B, h, w, C)).astype(np.float32)
    boxes = rng.uniform(-0.3, 1.3, (n, 4)).astype(np.float32)    # partly outside: extrapolated samples are 0
    boxes[0] = [0, 0, 1, 1]
    ind = rng.integers(0, B, n).astype(np.int32)
    cx, cy = orc.crop_and_resize(x, boxes, ind, (ph, pw)), orc.crop_and_resize(y, boxes, ind, (ph, pw))
    cz = orc.crop_and_resize((2 * x + y).astype(np.float32), boxes, ind, (ph, pw))
    assert np.allclose(cz, 2 * cx + cy, rtol=1e-4, atol=1e-4)
    if ph > 1 and pw > 1:                                        # the full box samples the corner pixels exactly
        assert np.array_equal(cx[0, 0, 0], x[ind[0], 0, 0]) and np.array_equal(cx[0, -1, -1], x[ind[0], -1, -1])
    g = rng.standard_normal(cx.shape).astype(np.float32)
    gx = orc.crop_and_resize_grad_image(g, boxes, ind, x.shape)
    lhs = float(np.sum(cx.astype(np.float64) * g))
    rhs = float(np.sum(x.astype(np.float64) * gx))
    assert abs(lhs - rhs) <= 1e-3 * (1 + abs(lhs))               # <crop(x), g> == <x, crop^T(g)>


@settings(max_examples=10, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10 ** 6), a=st.integers(50, 3000), p=st.sampled_from([1, 20, 300]))
def test_proposal_layer_output_rules(orc, seed, a, p):
    rng = np.random.default_rng(seed)
    B = 2
    anchors = np.ascontiguousarray(np.broadcast_to(random_boxes(rng, a, clusters=6), (B, a, 4)))
    fg = rng.uniform(0, 1, (B, a)).astype(np.float32)
    probs = np.stack([1 - fg, fg], -1).astype(np.float32)
    bbox = rng.standard_normal((B, a, 4)).astype(np.float32)
    r = orc.proposal_layer(probs, bbox, anchors, 600, p, SD, 0.7)
    k = min(600, a)
    for b in range(B):
        n = int(r["keep_count"][b])
        assert np.array_equal(r["topk_idx"][b], _rank_order(fg[b])[:k])                      # L:246
        assert np.all(r["proposals"][b, n:] == 0) and np.all(r["keep_idx"][b, n:] == -1)      # zero padding, L:229-230
        assert np.all((r["proposals"][b] >= 0) & (r["proposals"][b] <= 1))                    # clipped to [0,1], L:259
        assert np.array_equal(r["proposals"][b, :n], r["pre_nms_boxes"][b][r["keep_idx"][b, :n]])
        assert np.all(np.diff(r["keep_idx"][b, :n]) > 0)                                      # NMS keeps score order


@settings(max_examples=10, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 400), nc=st.sampled_from([2, 5, 81]),
       conf=st.sampled_from([0.0, 0.5, 0.9]))
def test_detection_layer_output_rules(orc, seed, n, nc, conf):
    from maskrcnn_tf2_b200 import synth
    rng = np.random.default_rng(seed)
    B, D = 2, 30
    rois = np.stack([random_boxes(rng, n, clusters=3) for _ in range(B)])
    z = 3 * rng.standard_normal((B, n, nc))
    probs = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(np.float32)
    deltas = rng.standard_normal((B, n, nc, 4)).astype(np.float32)
    meta = synth.image_meta(B, 512, nc)
    r = orc.detection_layer(rois, probs, deltas, meta, SD, conf, D, 0.3)
    for b in range(B):
        c = int(r["count"][b])
        det = r["detections"][b]
        assert np.all(det[c:] == 0)                                                           # L:498-500
        assert np.all(det[:c, 4] >= 1) and np.all(det[:c, 4] == np.round(det[:c, 4]))         # class > 0, stored as float
        assert np.all(np.diff(det[:c, 5]) <= 0)                                               # score order, L:486-490
        if conf:
            assert np.all(det[:c, 5] >= np.float32(conf))                                     # L:405
        assert np.all((det[:c, :4] >= 0) & (det[:c, :4] <= 1))                                # clipped to the window
        cls = probs[b].argmax(-1)
        sc = probs[b].max(-1)
        n_cand = int(np.sum((cls > 0) & ((sc >= np.float32(conf)) if conf else True)))
        assert c <= min(D, n_cand) and (c > 0) == (n_cand > 0)
