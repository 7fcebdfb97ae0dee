"""Differential tests that pin the CPU oracle against independent implementations available offline:
numpy stable sort, torchvision.ops.nms, torch grid_sample(align_corners=True) + autograd, and straight numpy
re-derivations of the layers that do not share code (or control flow) with oracle/mrcnn_oracle.c."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF
import torchvision

from conftest import random_boxes


def test_topk_matches_stable_argsort(orc):
    rng = np.random.default_rng(10)
    for n, k, quant in [(1000, 100, None), (5000, 5000, None), (4096, 600, 64), (300, 300, 4)]:
        s = rng.standard_normal(n).astype(np.float32)
        if quant:
            s = (np.round(s * quant) / quant).astype(np.float32)   # many exact ties
        ref = np.argsort(-s.astype(np.float64), kind="stable")[:k]
        assert np.array_equal(orc.topk(s, k), ref.astype(np.int32))


@pytest.mark.parametrize("clusters,thr", [(0, 0.5), (12, 0.7), (12, 0.3), (40, 0.5)])
def test_nms_matches_torchvision(orc, clusters, thr):
    rng = np.random.default_rng(11 + clusters)
    b = random_boxes(rng, 1500, clusters=clusters)
    s = rng.permutation(1500).astype(np.float32) / 1500.0           # tie-free scores
    ref = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
    got = orc.nms(b, s, 1500, thr)
    # torchvision tests inter/union > thr too, but may differ by one ulp on exact-threshold pairs: none here
    assert np.array_equal(got, ref.astype(np.int32))
    assert np.array_equal(orc.nms(b, s, 37, thr), ref[:37].astype(np.int32))


def _grid_for(boxes, H, W, ph, pw):
    """TF crop_and_resize sampling positions as a grid_sample(align_corners=True) grid."""
    nb = boxes.shape[0]
    ys = boxes[:, 0:1] * (H - 1) + np.arange(ph)[None, :] * ((boxes[:, 2:3] - boxes[:, 0:1]) * (H - 1) / (ph - 1))
    xs = boxes[:, 1:2] * (W - 1) + np.arange(pw)[None, :] * ((boxes[:, 3:4] - boxes[:, 1:2]) * (W - 1) / (pw - 1))
    gy = 2.0 * ys / (H - 1) - 1.0
    gx = 2.0 * xs / (W - 1) - 1.0
    grid = np.zeros((nb, ph, pw, 2), dtype=np.float64)
    grid[..., 0] = gx[:, None, :]
    grid[..., 1] = gy[:, :, None]
    return grid


def test_crop_and_resize_matches_grid_sample_and_autograd(orc):
    rng = np.random.default_rng(12)
    B, H, W, C, nb, ph, pw = 2, 13, 17, 5, 9, 7, 7
    img = rng.standard_normal((B, H, W, C)).astype(np.float32)
    y1 = rng.uniform(0, 0.5, nb); x1 = rng.uniform(0, 0.5, nb)
    boxes = np.stack([y1, x1, y1 + rng.uniform(0.1, 0.5, nb), x1 + rng.uniform(0.1, 0.5, nb)], 1).astype(np.float32)
    bi = rng.integers(0, B, nb).astype(np.int32)
    got = orc.crop_and_resize(img, boxes, bi, (ph, pw))
    timg = torch.from_numpy(img).double().permute(0, 3, 1, 2).requires_grad_(True)   # NCHW
    grid = torch.from_numpy(_grid_for(boxes.astype(np.float64), H, W, ph, pw))
    ref = TF.grid_sample(timg[torch.from_numpy(bi).long()], grid, mode="bilinear", padding_mode="zeros",
                         align_corners=True)                                          # [nb,C,ph,pw]
    assert np.allclose(got, ref.detach().permute(0, 2, 3, 1).numpy(), rtol=1e-5, atol=1e-5)
    g = rng.standard_normal((nb, ph, pw, C)).astype(np.float32)
    ref.backward(torch.from_numpy(g).double().permute(0, 3, 1, 2))
    gi = orc.crop_and_resize_grad_image(g, boxes, bi, (B, H, W, C))
    assert np.allclose(gi, timg.grad.permute(0, 2, 3, 1).numpy(), rtol=1e-4, atol=1e-5)


def _levels_float64(boxes, img_area):
    h = (boxes[..., 2] - boxes[..., 0]).astype(np.float64)
    w = (boxes[..., 3] - boxes[..., 1]).astype(np.float64)
    lv = np.log2(np.sqrt(h * w) / (244.0 / np.sqrt(img_area)))
    assert np.all(np.abs(lv - np.floor(lv) - 0.5) > 1e-4)          # away from rounding boundaries
    return np.clip(4 + np.rint(lv).astype(np.int64), 2, 5)


def test_pyramid_roi_align_literal_flow_equals_direct_crops(orc):
    rng = np.random.default_rng(13)
    B, N, C, ph, pw = 3, 40, 8, 7, 7
    side = np.exp(rng.uniform(np.log(20), np.log(800), (B, N))) / 1024.0
    ar = np.exp(rng.uniform(-0.5, 0.5, (B, N)))
    h, w = side * ar, side / ar
    y1, x1 = rng.uniform(0, 1 - np.minimum(h, 0.99), (B, N)), rng.uniform(0, 1 - np.minimum(w, 0.99), (B, N))
    boxes = np.stack([y1, x1, y1 + h, x1 + w], -1).astype(np.float32)
    fm = [rng.standard_normal((B, s, s, C)).astype(np.float32) for s in (64, 32, 16, 8)]
    r = orc.pyramid_roi_align(boxes, 1024.0, 1024.0, fm, (ph, pw))
    lv = _levels_float64(boxes, 1024.0 * 1024.0)
    assert np.array_equal(r["level"], lv)
    order = []
    for v in lv.reshape(-1):
        if v not in order:
            order.append(int(v))
    table = {v: i for i, v in enumerate(order)}
    for b in range(B):
        for n in range(N):
            m = table[int(lv[b, n])]
            assert r["roi_map"][b, n] == m
            ref = orc.crop_and_resize(fm[m], [boxes[b, n]], [b], (ph, pw))[0]
            assert np.array_equal(r["out"][b, n], ref)
    # gradient of the literal flow = per-map scatter of the same ROIs
    g = rng.standard_normal((B, N, ph, pw, C)).astype(np.float32)
    grads = orc.pyramid_roi_align_grad(g, boxes, 1024.0, 1024.0, [f.shape for f in fm])
    for m in range(4):
        sel = [(b, n) for b in range(B) for n in range(N) if table[int(lv[b, n])] == m]
        ref = np.zeros_like(fm[m])
        if sel:
            ref = orc.crop_and_resize_grad_image(np.stack([g[b, n] for b, n in sel]),
                                                 np.stack([boxes[b, n] for b, n in sel]),
                                                 np.array([b for b, _ in sel], np.int32), fm[m].shape)
        assert np.array_equal(grads[m], ref)


def _decode_np(boxes, deltas):
    b, d = boxes.astype(np.float32), deltas.astype(np.float32)
    h = b[:, 2] - b[:, 0]; w = b[:, 3] - b[:, 1]
    cy = b[:, 0] + np.float32(0.5) * h; cx = b[:, 1] + np.float32(0.5) * w
    cy = cy + d[:, 0] * h; cx = cx + d[:, 1] * w
    h = h * np.exp(d[:, 2]); w = w * np.exp(d[:, 3])
    y1 = cy - np.float32(0.5) * h; x1 = cx - np.float32(0.5) * w
    return np.stack([y1, x1, y1 + h, x1 + w], 1).astype(np.float32)


def test_proposal_layer_against_numpy_pipeline(orc):
    rng = np.random.default_rng(14)
    B, A, K, P = 2, 3000, 500, 120
    probs = rng.uniform(0, 1, (B, A, 2)).astype(np.float32)
    bbox = (0.5 * rng.standard_normal((B, A, 4))).astype(np.float32)
    anchors = np.broadcast_to(random_boxes(rng, A, clusters=10), (B, A, 4)).copy()
    sd = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
    r = orc.proposal_layer(probs, bbox, anchors, K, P, sd, 0.7)
    for b in range(B):
        ix = np.argsort(-probs[b, :, 1].astype(np.float64), kind="stable")[:K]
        assert np.array_equal(r["topk_idx"][b], ix.astype(np.int32))
        boxes = np.clip(_decode_np(anchors[b, ix], bbox[b, ix] * sd), 0, 1)
        assert np.allclose(r["pre_nms_boxes"][b], boxes, rtol=1e-5, atol=1e-6)      # numpy exp vs orc_expf: ulps
        keep = torchvision.ops.nms(torch.from_numpy(r["pre_nms_boxes"][b]),
                                   torch.from_numpy(probs[b, ix, 1].copy()), 0.7).numpy()[:P]
        n = r["keep_count"][b]
        assert n == len(keep) and np.array_equal(r["keep_idx"][b, :n], keep)
        assert np.array_equal(r["proposals"][b, :n], r["pre_nms_boxes"][b][keep])
        assert np.array_equal(r["proposals"][b, n:], np.zeros((P - n, 4), np.float32))


def test_proposal_layer_gradient_against_torch_autograd(orc):
    # Q7: the reference back-propagates through gather -> clip -> decode -> std-dev scale into rpn_bbox
    rng = np.random.default_rng(17)
    B, A, K, P = 2, 900, 300, 80
    probs = rng.uniform(0, 1, (B, A, 2)).astype(np.float32)
    bbox = (0.8 * rng.standard_normal((B, A, 4))).astype(np.float32)
    anchors = np.stack([random_boxes(rng, A, clusters=6) for _ in range(B)])
    sd = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
    r = orc.proposal_layer(probs, bbox, anchors, K, P, sd, 0.7)
    g = rng.standard_normal((B, P, 4)).astype(np.float32)
    got = orc.proposal_layer_grad(g, bbox, anchors, r["topk_idx"], r["keep_idx"], sd)
    tb = torch.tensor(bbox, dtype=torch.float64, requires_grad=True)
    ta = torch.tensor(anchors, dtype=torch.float64)
    tsd = torch.tensor(sd, dtype=torch.float64)
    outs = []
    for b in range(B):
        ix = torch.tensor(r["topk_idx"][b]).long()
        d, an = tb[b][ix] * tsd, ta[b][ix]
        h, w = an[:, 2] - an[:, 0], an[:, 3] - an[:, 1]
        cy, cx = an[:, 0] + 0.5 * h + d[:, 0] * h, an[:, 1] + 0.5 * w + d[:, 1] * w
        h2, w2 = h * torch.exp(d[:, 2]), w * torch.exp(d[:, 3])
        y1, x1 = cy - 0.5 * h2, cx - 0.5 * w2
        boxes = torch.stack([y1, x1, y1 + h2, x1 + w2], 1).clamp(0, 1)
        n = int(r["keep_count"][b])
        o = torch.zeros(P, 4, dtype=torch.float64)
        o[:n] = boxes[torch.tensor(r["keep_idx"][b][:n]).long()]
        outs.append(o)
    torch.stack(outs).backward(torch.tensor(g, dtype=torch.float64))
    ref = tb.grad.numpy()
    assert (np.abs(ref).sum(-1) > 0).sum() > 50 and (ref == 0).mean() > 0.5      # sparse scatter, some coords clipped
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-6)


def test_detection_layer_against_numpy_pipeline(orc):
    rng = np.random.default_rng(15)
    B, N, NC, D = 2, 400, 11, 50
    rois = np.stack([random_boxes(rng, N, clusters=8) for _ in range(B)])
    logits = 3.0 * rng.standard_normal((B, N, NC))
    probs = (np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)).astype(np.float32)
    deltas = rng.standard_normal((B, N, NC, 4)).astype(np.float32) * 0.3
    meta = np.zeros((B, 12 + NC), np.float32)
    meta[:, 4:7] = (512, 512, 3)
    meta[0, 7:11] = (0, 0, 512, 512)
    meta[1, 7:11] = (32, 64, 480, 448)
    sd = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
    r = orc.detection_layer(rois, probs, deltas, meta, sd, 0.5, D, 0.3)
    for b in range(B):
        cls = probs[b].argmax(1)
        sc = probs[b][np.arange(N), cls]
        win = ((meta[b, 7:11] - np.array([0, 0, 1, 1], np.float32)) / np.float32(511.0)).astype(np.float32)
        boxes = _decode_np(rois[b], deltas[b][np.arange(N), cls] * sd)
        boxes = np.stack([np.clip(boxes[:, 0], win[0], win[2]), np.clip(boxes[:, 1], win[1], win[3]),
                          np.clip(boxes[:, 2], win[0], win[2]), np.clip(boxes[:, 3], win[1], win[3])], 1)
        keep = np.where((cls > 0) & (sc >= np.float32(0.5)))[0]
        # detections are compared on the oracle's own refined boxes (exp ulps), selection logic independently
        sel = torchvision.ops.nms(torch.from_numpy(boxes[keep]), torch.from_numpy(sc[keep]), 0.3).numpy()[:D]
        n = r["count"][b]
        assert n == len(sel)
        got = r["detections"][b]
        assert np.allclose(got[:n, :4], boxes[keep][sel], rtol=1e-5, atol=1e-6)
        assert np.array_equal(got[:n, 4], cls[keep][sel].astype(np.float32))
        assert np.array_equal(got[:n, 5], sc[keep][sel])
        assert np.array_equal(got[n:], np.zeros((D - n, 6), np.float32))


def test_detection_target_layer_against_numpy_pipeline(orc):
    rng = np.random.default_rng(16)
    B, P, G, T, MH = 2, 300, 12, 60, 64
    props = np.stack([random_boxes(rng, P, min_size=0.05, max_size=0.4, clusters=6) for _ in range(B)])
    props[:, -20:] = 0.0                                                     # zero padding rows
    gtb = np.zeros((B, G, 4), np.float32)
    gtc = np.zeros((B, G), np.int32)
    for b in range(B):
        gtb[b, :6] = props[b, rng.choice(P - 20, 6, replace=False)]          # GT boxes coincide with proposals
        gtc[b, :6] = rng.integers(1, 10, 6)
    gtc[1, 5] = -3                                                           # one crowd box
    masks = (rng.uniform(0, 1, (B, MH, MH, G)) < 0.5).astype(np.uint8)
    keys = rng.integers(0, 2 ** 32, (B, P), dtype=np.uint64).astype(np.uint32)
    sd = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
    r = orc.detection_target_layer(props, gtc, gtb, masks, keys, T, 0.33, sd, (28, 28))
    for b in range(B):
        nz = np.abs(props[b]).sum(1) != 0
        gt_ok = np.where(gtc[b] > 0)[0]
        cr = np.where(gtc[b] < 0)[0]
        def iou(a, c):
            y1 = np.maximum(a[:, None, 0], c[None, :, 0]); x1 = np.maximum(a[:, None, 1], c[None, :, 1])
            y2 = np.minimum(a[:, None, 2], c[None, :, 2]); x2 = np.minimum(a[:, None, 3], c[None, :, 3])
            inter = np.maximum(x2 - x1, 0) * np.maximum(y2 - y1, 0)
            aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ac = (c[:, 2] - c[:, 0]) * (c[:, 3] - c[:, 1])
            return inter / (aa[:, None] + ac[None, :] - inter)
        ov = iou(props[b], gtb[b, gt_ok])
        mx = ov.max(1)
        crowd_ok = (iou(props[b], gtb[b, cr]).max(1) < 0.001) if len(cr) else np.ones(P, bool)
        pos = np.where(nz & (mx >= 0.5))[0]
        neg = np.where(nz & (mx < 0.5) & crowd_ok)[0]
        pos = pos[np.lexsort((pos, keys[b, pos]))][:int(T * 0.33)]
        ncount = int(np.float32(1.0 / 0.33) * np.float32(len(pos))) - len(pos)
        neg = neg[np.lexsort((neg, keys[b, neg]))][:ncount]
        assert r["counts"][b].tolist() == [len(pos), len(neg)]
        assert np.array_equal(r["rois"][b, :len(pos)], props[b, pos])
        assert np.array_equal(r["rois"][b, len(pos):len(pos) + len(neg)], props[b, neg])
        assert np.array_equal(r["rois"][b, len(pos) + len(neg):], np.zeros((T - len(pos) - len(neg), 4), np.float32))
        assign = gt_ok[ov[pos].argmax(1)]
        assert np.array_equal(r["class_ids"][b, :len(pos)], gtc[b, assign])
        assert np.all(r["class_ids"][b, len(pos):] == 0)
        g, p = gtb[b, assign].astype(np.float64), props[b, pos].astype(np.float64)
        ph_, pw_ = p[:, 2] - p[:, 0], p[:, 3] - p[:, 1]
        gh, gw = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1]
        ref = np.stack([((g[:, 0] + gh / 2) - (p[:, 0] + ph_ / 2)) / ph_ / 0.1,
                        ((g[:, 1] + gw / 2) - (p[:, 1] + pw_ / 2)) / pw_ / 0.1,
                        np.log(gh / (ph_ + 1e-3)) / 0.2, np.log(gw / (pw_ + 1e-3)) / 0.2], 1)
        assert np.allclose(r["deltas"][b, :len(pos)], ref, rtol=1e-4, atol=1e-4)
        assert np.all(r["deltas"][b, len(pos):] == 0)
        for s_, (i, ga) in enumerate(zip(pos, assign)):
            m = masks[b, :, :, ga].astype(np.float32)[None, :, :, None]
            ref_m = np.rint(orc.crop_and_resize(m, [props[b, i]], [0], (28, 28))[0, :, :, 0])
            assert np.array_equal(r["masks"][b, s_], ref_m)
        assert np.all(r["masks"][b, len(pos):] == 0)
