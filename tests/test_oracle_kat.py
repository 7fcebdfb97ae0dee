"""Known-answer tests that pin the CPU oracle: every expected value below is derived by hand from the TF kernel
definitions (TopKV2, NonMaxSuppressionV3, CropAndResize[GradImage]) and the reference's control flow."""
import numpy as np
import pytest


def test_exp_log_within_one_ulp(orc):
    rng = np.random.default_rng(0)
    x = rng.uniform(-30, 30, 4000).astype(np.float32)
    e = orc.expf(x).astype(np.float64)
    t = np.exp(x.astype(np.float64))
    assert np.max(np.abs(e - t) / np.spacing(t.astype(np.float32)).astype(np.float64)) <= 1.0
    x = np.exp(rng.uniform(-20, 20, 4000)).astype(np.float32)
    l = orc.logf(x).astype(np.float64)
    t = np.log(x.astype(np.float64))
    assert np.max(np.abs(l - t) / np.spacing(np.abs(t).astype(np.float32)).astype(np.float64)) <= 1.0
    assert orc.expf([0.0])[0] == 1.0 and orc.logf([1.0])[0] == 0.0
    assert np.isinf(orc.expf([89.0])[0]) and orc.expf([-104.0])[0] == 0.0
    assert orc.logf([0.0])[0] == -np.inf and np.isnan(orc.logf([-1.0])[0])


def test_topk_ties_lower_index_first(orc):
    s = np.array([0.5, 0.9, 0.5, 0.9, 0.1, 0.5], dtype=np.float32)
    assert orc.topk(s, 6).tolist() == [1, 3, 0, 2, 5, 4]
    assert orc.topk(s, 3).tolist() == [1, 3, 0]
    assert orc.topk(np.zeros(5, np.float32), 3).tolist() == [0, 1, 2]


def test_nms_three_boxes(orc):
    # box1 overlaps box0 with IoU 0.81 -> suppressed at 0.5; box2 disjoint
    b = np.array([[0, 0, 1, 1], [0, 0, 0.9, 0.9], [2, 2, 3, 3]], dtype=np.float32)
    s = np.array([0.9, 0.8, 0.7], dtype=np.float32)
    assert orc.nms(b, s, 3, 0.5).tolist() == [0, 2]
    assert orc.nms(b, s, 3, 0.85).tolist() == [0, 1, 2]      # 0.81 is not > 0.85
    assert orc.nms(b, s, 1, 0.5).tolist() == [0]             # max_output_size
    assert abs(orc.tf_iou(b, 0, 1) - 0.81) < 1e-6


def test_nms_strict_threshold_and_order(orc):
    # IoU exactly 0.5: two unit-height boxes sharing 2/3 of the width: inter=2, union=4
    b = np.array([[0, 0, 1, 3], [0, 1, 1, 4]], dtype=np.float32)
    assert orc.tf_iou(b, 0, 1) == 0.5
    s = np.array([0.3, 0.3], dtype=np.float32)
    assert orc.nms(b, s, 2, 0.5).tolist() == [0, 1]          # kept: IoU is not > thr; tie -> lower index first
    assert orc.nms(b, s, 2, 0.49).tolist() == [0]


def test_nms_zero_area_flipped_and_neg_inf(orc):
    b = np.array([[0, 0, 1, 1],      # 0
                  [1, 1, 0, 0],      # 1: flipped corners = same box as 0 after min/max normalisation
                  [0.5, 0.5, 0.5, 0.9],  # 2: zero area: IoU 0 with everything, still selected
                  [0.5, 0.5, 0.5, 0.9],  # 3: duplicate zero-area box: not suppressed either
                  [0, 0, 1, 1]],     # 4: score -inf -> never a candidate
                 dtype=np.float32)
    s = np.array([0.9, 0.8, 0.7, 0.6, -np.inf], dtype=np.float32)
    assert orc.nms(b, s, 5, 0.5).tolist() == [0, 2, 3]
    s2 = np.array([0.9, np.nan, 0.7, 0.6, 0.1], dtype=np.float32)
    assert orc.nms(b, s2, 5, 0.5).tolist() == [0, 2, 3]      # NaN never passes score > -inf; 4 suppressed by 0


def test_apply_deltas_and_clip(orc):
    box = np.array([[0.25, 0.25, 0.75, 0.75]], dtype=np.float32)
    assert np.array_equal(orc.apply_box_deltas(box, np.zeros((1, 4), np.float32)), box)
    out = orc.apply_box_deltas(box, np.array([[0.5, -0.5, 0.0, 0.0]], np.float32))
    assert np.allclose(out, [[0.5, 0.0, 1.0, 0.5]], atol=1e-7)
    out = orc.apply_box_deltas(box, np.array([[0, 0, np.log(2.0), 0]], np.float32))
    assert np.allclose(out, [[0.0, 0.25, 1.0, 0.75]], atol=1e-6)
    c = orc.clip_boxes(np.array([[-0.5, 0.2, 1.5, 0.8]], np.float32), [0, 0, 1, 1])
    assert np.array_equal(c, np.array([[0, 0.2, 1, 0.8]], np.float32))


def test_crop_and_resize_known_values(orc):
    # image value = 10*y + x on a 4x4 grid, one channel
    img = (10 * np.arange(4)[:, None] + np.arange(4)[None, :]).astype(np.float32).reshape(1, 4, 4, 1)
    full = orc.crop_and_resize(img, [[0, 0, 1, 1]], [0], (4, 4))
    assert np.array_equal(full[0, :, :, 0], img[0, :, :, 0])           # identity crop
    two = orc.crop_and_resize(img, [[0, 0, 1, 1]], [0], (2, 2))
    assert np.array_equal(two[0, :, :, 0], [[0, 3], [30, 33]])         # corners only
    mid = orc.crop_and_resize(img, [[0.5, 0.5, 0.5, 0.5]], [0], (2, 2))
    assert np.allclose(mid[0, :, :, 0], 16.5)                          # in = 1.5 both axes: bilinear centre
    one = orc.crop_and_resize(img, [[0, 0, 1, 1]], [0], (1, 1))
    assert np.allclose(one[0, 0, 0, 0], 16.5)                          # crop size 1 samples the box centre
    out = orc.crop_and_resize(img, [[-1, 0, 1, 1]], [0], (3, 2))       # first row at in_y = -3 -> extrapolated 0
    assert np.array_equal(out[0, 0, :, 0], [0, 0]) and np.array_equal(out[0, 1, :, 0], [0, 3])
    assert np.array_equal(out[0, 2, :, 0], [30, 33])


def test_crop_and_resize_grad_known_values(orc):
    g = np.ones((1, 2, 2, 1), dtype=np.float32)
    # box = full image on a 2x2 map: every output pixel hits exactly one input pixel with weight 1
    gi = orc.crop_and_resize_grad_image(g, [[0, 0, 1, 1]], [0], (1, 2, 2, 1))
    assert np.array_equal(gi[0, :, :, 0], [[1, 1], [1, 1]])
    # all four samples at the centre (0.5,0.5): each spreads 1/4 to every corner
    gi = orc.crop_and_resize_grad_image(g, [[0.5, 0.5, 0.5, 0.5]], [0], (1, 2, 2, 1))
    assert np.allclose(gi[0, :, :, 0], 1.0)
    # out-of-range rows contribute nothing
    gi = orc.crop_and_resize_grad_image(g, [[-2, 0, -1, 1]], [0], (1, 2, 2, 1))
    assert np.array_equal(gi, np.zeros_like(gi))


@pytest.mark.parametrize("side_px,expected", [(40, 2), (86, 2), (87, 3), (172, 3), (173, 4), (345, 4), (346, 5),
                                              (900, 5), (0, 2)])
def test_roi_level_boundaries_use_244(orc, side_px, expected):
    # quirk Q1: boundaries at 244 * 2^(k-0.5) / ... = 86.3 / 172.5 / 345.1 px for a 1024^2 image
    s = side_px / 1024.0
    assert orc.roi_level([0.1, 0.1, 0.1 + s, 0.1 + s], 1024.0, 1024.0) == expected


def test_roi_level_degenerate_boxes_go_to_level_2(orc):
    assert orc.roi_level([0, 0, 0, 0], 1024.0, 1024.0) == 2            # log(0) = -inf -> INT_MIN -> clamp
    assert orc.roi_level([0.5, 0.5, 0.4, 0.9], 1024.0, 1024.0) == 2    # negative area -> NaN -> INT_MIN


def test_pyramid_first_appearance_map_table(orc):
    # levels in flattened order: 4, 2, 4, 5 -> unique = [4, 2, 5] -> level 4 reads P2 (index 0), level 2 reads
    # P3 (index 1), level 5 reads P4 (index 2) (quirk Q2)
    def box(side_px):
        s = side_px / 1024.0
        return [0.1, 0.1, 0.1 + s, 0.1 + s]
    boxes = np.array([[box(300), box(40)], [box(200), box(800)]], dtype=np.float32)
    rng = np.random.default_rng(3)
    fm = [rng.standard_normal((2, h, h, 4)).astype(np.float32) for h in (16, 8, 4, 2)]
    r = orc.pyramid_roi_align(boxes, 1024.0, 1024.0, fm, (2, 2))
    assert r["level"].tolist() == [[4, 2], [4, 5]]
    assert r["roi_map"].tolist() == [[0, 1], [0, 2]]
    # each ROI equals a direct crop from the mapped feature map of its own image
    for b in range(2):
        for n in range(2):
            m = r["roi_map"][b, n]
            ref = orc.crop_and_resize(fm[m], [boxes[b, n]], [b], (2, 2))[0]
            assert np.array_equal(r["out"][b, n], ref)
    canon = orc.pyramid_roi_align(boxes, 1024.0, 1024.0, fm, (2, 2), map_mode=1)
    assert canon["roi_map"].tolist() == [[2, 0], [2, 3]]


def test_pyramid_zero_padded_rois_sample_pixel_00(orc):
    boxes = np.zeros((1, 3, 4), dtype=np.float32)
    boxes[0, 0] = [0.2, 0.2, 0.6, 0.6]
    rng = np.random.default_rng(4)
    fm = [rng.standard_normal((1, h, h, 4)).astype(np.float32) for h in (16, 8, 4, 2)]
    r = orc.pyramid_roi_align(boxes, 1024.0, 1024.0, fm, (3, 3))
    # box 0 is level 5 (0.4*1024 = 410 px) and appears first -> map 0; padded rows are level 2 -> map 1 (quirk Q5)
    assert r["roi_map"].tolist() == [[0, 1, 1]]
    assert np.array_equal(r["out"][0, 1], np.broadcast_to(fm[1][0, 0, 0], (3, 3, 4)))


def test_detection_layer_hand_case(orc):
    # 4 ROIs, 3 classes.  ROI0: class 1 @0.9; ROI1: class 1 @0.8 overlapping ROI0 (suppressed, IoU>0.3 although a
    # "different image region"); ROI2: background; ROI3: class 2 @0.6 < min_conf.  Class-agnostic NMS (Q3): a
    # class-2 ROI4 @0.95 overlapping ROI0 suppresses it even though the classes differ.
    rois = np.array([[[0.1, 0.1, 0.5, 0.5], [0.12, 0.12, 0.5, 0.5], [0.6, 0.6, 0.9, 0.9], [0.6, 0.1, 0.9, 0.4],
                      [0.1, 0.1, 0.5, 0.52]]], dtype=np.float32)
    probs = np.array([[[0.05, 0.9, 0.05], [0.1, 0.8, 0.1], [0.8, 0.1, 0.1], [0.2, 0.2, 0.6], [0.02, 0.03, 0.95]]],
                     dtype=np.float32)
    deltas = np.zeros((1, 5, 3, 4), dtype=np.float32)
    meta = np.zeros((1, 15), dtype=np.float32)
    meta[0, 4:7] = (1024, 1024, 3)
    meta[0, 7:11] = (0, 0, 1024, 1024)
    r = orc.detection_layer(rois, probs, deltas, meta, [0.1, 0.1, 0.2, 0.2], 0.7, 4, 0.3)
    assert r["count"].tolist() == [1]
    d = r["detections"][0]
    assert np.allclose(d[0], [0.1, 0.1, 0.5, 0.52, 2.0, 0.95], atol=1e-6)
    assert np.array_equal(d[1:], np.zeros((3, 6), np.float32))
    # window clipping: window (0,0)-(512,1024) px of a 1024 image -> y clipped to 511/1023
    meta[0, 7:11] = (0, 0, 512, 1024)
    r = orc.detection_layer(rois, probs, deltas, meta, [0.1, 0.1, 0.2, 0.2], 0.0, 4, 0.3)
    # min_conf falsy: filter skipped (L:404).  Order ROI4(.95), ROI0, ROI1 (both suppressed by ROI4), ROI3(.6):
    # ROI3 lies below the window, is clipped to zero height, has IoU 0 with everything and is kept.
    assert r["count"].tolist() == [2]
    assert np.all(r["detections"][0, :2, 2] <= np.float32(511.0 / 1023.0))
    assert r["detections"][0, 1, 4] == 2.0 and r["detections"][0, 1, 0] == r["detections"][0, 1, 2]


def test_detection_target_hand_case(orc):
    # one GT box; proposal 0 identical (IoU 1, positive), proposal 1 disjoint (negative), proposal 2 zero padding
    P, G, T = 4, 2, 4
    props = np.zeros((1, P, 4), np.float32)
    props[0, 0] = [0.2, 0.2, 0.6, 0.6]
    props[0, 1] = [0.7, 0.7, 0.9, 0.9]
    props[0, 3] = [0.2, 0.2, 0.6, 0.5]        # IoU 0.75 with the GT: positive too
    gtb = np.zeros((1, G, 4), np.float32)
    gtb[0, 0] = [0.2, 0.2, 0.6, 0.6]
    gtc = np.array([[7, 0]], np.int32)
    masks = np.zeros((1, 16, 16, G), np.uint8)
    masks[0, :, :, 0] = 1
    keys = np.array([[5, 1, 0, 2]], np.uint32)   # positives ordered by key: row 3 (key 2) then row 0 (key 5)
    r = orc.detection_target_layer(props, gtc, gtb, masks, keys, T, 0.5, [0.1, 0.1, 0.2, 0.2], (4, 4))
    assert r["counts"].tolist() == [[2, 1]]      # positive cap int(4*0.5)=2; negatives int(2.0*2)-2 = 2 -> only 1 exists
    assert np.array_equal(r["rois"][0, 0], props[0, 3]) and np.array_equal(r["rois"][0, 1], props[0, 0])
    assert np.array_equal(r["rois"][0, 2], props[0, 1]) and np.array_equal(r["rois"][0, 3], [0, 0, 0, 0])
    assert r["class_ids"].tolist() == [[7, 7, 0, 0]]
    assert np.allclose(r["deltas"][0, 1, :2], 0.0) and r["deltas"][0, 1, 2] < 0   # log(h/(h+1e-3)) slightly < 0
    assert np.allclose(r["deltas"][0, 0, 1], (0.4 - 0.35) / 0.3 / 0.1, rtol=1e-5)
    assert np.array_equal(r["masks"][0, 0], np.ones((4, 4))) and np.array_equal(r["masks"][0, 2], np.zeros((4, 4)))


def test_rpn_softmax_known_answers(orc):
    """Keras softmax over two logits in TF's order exp(l - max) * (1 / sum): exact halves, saturation, invariance to a
    common offset, and agreement with float64 softmax to fp32 rounding."""
    x = np.array([[0, 0], [3, 3], [100, -100], [-100, 100], [1, 2], [1001, 1002]], np.float32)
    p = orc.rpn_softmax(x)
    assert np.array_equal(p[0], [0.5, 0.5]) and np.array_equal(p[1], [0.5, 0.5])
    assert np.array_equal(p[2], [1.0, 0.0]) and np.array_equal(p[3], [0.0, 1.0])
    assert np.array_equal(p[4], p[5])                                       # only the difference matters
    rng = np.random.default_rng(3)
    y = rng.normal(0, 4, (4096, 2)).astype(np.float32)
    e = np.exp(y.astype(np.float64) - y.max(1, keepdims=True))
    assert np.abs(orc.rpn_softmax(y) - e / e.sum(1, keepdims=True)).max() < 3e-7
