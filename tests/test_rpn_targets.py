"""build_rpn_targets (utils.py:154-262): the oracle and the CUDA path against outputs of the REFERENCE'S OWN numpy code.

tests/golden/reference_rpn_targets_golden.npz was written by tests/golden/make_reference_rpn_golden.py, which imports
/root/reference/src/common/utils.py in the build container and runs its build_rpn_targets with np.random.choice
replaced by the key-driven rule (keep the largest keys, ties -> lower anchor index).  This row of the path is therefore
PINNED: rpn_match is compared bit-exactly (int), rpn_bbox's linear columns bit-exactly (float64) and its log columns
to 4 ulp of float64 (CUDA's log() is not glibc's)."""
import hashlib
import os

import numpy as np
import pytest

from maskrcnn_tf2_b200 import synth

SD = (0.1, 0.1, 0.2, 0.2)
LOG_RTOL = 4 * 2.0 ** -52


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_rpn_targets_golden.npz"))


def keys_for(seed, b, A):  # the generator script's key stream
    return np.random.default_rng(seed * 1000 + b).random(A, dtype=np.float32)


def case(G, tag):
    S, R, seed = int(G[tag + "_S"]), int(G[tag + "_R"]), int(G[tag + "_key_seed"])
    an = synth.pyramid_anchors_px(S)
    cls, box = G[tag + "_gt_class_ids"], G[tag + "_gt_boxes"]
    keys = np.stack([keys_for(seed, b, an.shape[0]) for b in range(cls.shape[0])])
    return an, cls, box, keys, R


def check_bbox(got, want):
    assert np.array_equal(got[..., :2], want[..., :2])
    np.testing.assert_allclose(got[..., 2:], want[..., 2:], rtol=LOG_RTOL, atol=0)


def check_small(G, match, bbox):
    assert np.array_equal(match, G["small_rpn_match"].astype(np.int32))
    check_bbox(bbox, G["small_rpn_bbox"])


def check_coco(G, match, bbox):
    for b in range(match.shape[0]):
        dig = np.frombuffer(hashlib.sha256(np.ascontiguousarray(match[b], np.int32).tobytes()).digest(), np.uint8)
        pos, neg = G["coco_pos_idx"][b], G["coco_neg_idx"][b]
        assert np.array_equal(np.where(match[b] == 1)[0], pos[pos >= 0])
        assert np.array_equal(np.where(match[b] == -1)[0], neg[neg >= 0])
        assert np.array_equal(dig, G["coco_rpn_match_sha256"][b])
    check_bbox(bbox, G["coco_rpn_bbox"])


# ---- CPU: the oracle against the reference's outputs -----------------------------------------------------------

def test_oracle_matches_reference_small(G, orc):
    an, cls, box, keys, R = case(G, "small")
    r = orc.build_rpn_targets(an, cls, box, keys, R, SD)
    check_small(G, r["rpn_match"], r["rpn_bbox"])
    # the matching rule alone (no subsampling): R = 2A
    ru = orc.build_rpn_targets(an, cls, box, keys, 2 * an.shape[0], SD)
    assert np.array_equal(ru["rpn_match"], G["small_rpn_match_unsampled"].astype(np.int32))
    # edge cases the fixture holds: zero-area instance -> every anchor positive; crowd over the whole image -> no
    # negatives; a tiny instance matched only through the forced `overlaps == max` rule
    assert (G["small_rpn_match_unsampled"][3] == 1).all()
    assert r["counts"].tolist()[3] == [R // 2, 0] and r["counts"].tolist()[5][1] == 0
    assert r["counts"].tolist()[4] == [R // 2, R // 2]


def test_oracle_matches_reference_coco_shape(G, orc):
    an, cls, box, keys, R = case(G, "coco")
    r = orc.build_rpn_targets(an, cls, box, keys, R, SD)
    check_coco(G, r["rpn_match"], r["rpn_bbox"])


def test_oracle_subsampling_properties(orc):
    rng = np.random.default_rng(5)
    S, R, B, Gp = 256, 32, 4, 12
    an = synth.pyramid_anchors_px(S)
    A = an.shape[0]
    cls = np.zeros((B, Gp), np.int32)
    box = np.zeros((B, Gp, 4), np.int32)
    for b in range(B):
        n = 3 + 2 * b
        y1, x1 = rng.integers(0, S - 40, n), rng.integers(0, S - 40, n)
        box[b, :n] = np.stack([y1, x1, y1 + rng.integers(8, 120, n), x1 + rng.integers(8, 120, n)], 1)
        cls[b, :n] = rng.integers(1, 80, n)
    keys = rng.random((B, A), dtype=np.float32)
    full = orc.build_rpn_targets(an, cls, box, keys, 2 * A, SD)["rpn_match"]
    r = orc.build_rpn_targets(an, cls, box, keys, R, SD)
    m = r["rpn_match"]
    for b in range(B):
        assert ((m[b] != 0) <= (m[b] == full[b])).all()                 # only resets to neutral
        npos, nneg = int((m[b] == 1).sum()), int((m[b] == -1).sum())
        assert npos == min(int((full[b] == 1).sum()), R // 2) and npos + nneg <= R
        assert r["counts"][b].tolist() == [npos, nneg]
        kept = keys[b][m[b] == -1]                                       # the kept negatives carry the largest keys
        dropped = keys[b][(full[b] == -1) & (m[b] == 0)]
        assert dropped.size == 0 or kept.min() >= dropped.max()
        assert not r["rpn_bbox"][b, npos:].any()


# ---- GPU: the CUDA path against the reference's outputs and the oracle -----------------------------------------

def _cuda_targets(dev, an, cls, box, keys, R, **kw):
    import torch
    from maskrcnn_tf2_b200 import functional as F
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = F.rpn_targets_forward(t(an), t(cls.astype(np.int32)), t(box.astype(np.int32)), t(keys), R, SD, **kw)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


@pytest.mark.gpu
def test_cuda_matches_reference_small(G, dev):
    an, cls, box, keys, R = case(G, "small")
    match, bbox, bbox32, counts = _cuda_targets(dev, an, cls, box, keys, R, return_f32=True, return_counts=True)
    check_small(G, match, bbox)
    assert np.array_equal(counts[:, 0], (match == 1).sum(1)) and np.array_equal(counts[:, 1], (match == -1).sum(1))
    assert np.array_equal(bbox32, bbox.astype(np.float32))
    match_all, _ = _cuda_targets(dev, an, cls, box, keys, 2 * an.shape[0])
    assert np.array_equal(match_all, G["small_rpn_match_unsampled"].astype(np.int32))


@pytest.mark.gpu
def test_cuda_matches_reference_coco_shape(G, dev):
    an, cls, box, keys, R = case(G, "coco")
    match, bbox = _cuda_targets(dev, an, cls, box, keys, R)
    check_coco(G, match, bbox)


@pytest.mark.gpu
@pytest.mark.parametrize("S,B,Gp,R,n_real,crowds", [(1024, 8, 100, 256, 20, 1), (512, 5, 64, 256, 40, 3),
                                                     (256, 3, 7, 16, 5, 0), (1024, 2, 1024, 512, 300, 10)])
def test_cuda_matches_oracle_batches(orc, dev, S, B, Gp, R, n_real, crowds):
    rng = np.random.default_rng(S + B)
    an = synth.pyramid_anchors_px(S)
    A = an.shape[0]
    cls = np.zeros((B, Gp), np.int32)
    box = np.zeros((B, Gp, 4), np.int32)
    for b in range(B):
        rows = rng.permutation(Gp)[:n_real]                              # real rows scattered between padding rows
        side = rng.uniform(S / 40, S / 2, (n_real, 2))
        c = rng.uniform(0, S, (n_real, 2))
        y1, x1 = np.clip(c[:, 0] - side[:, 0] / 2, 0, S - 2), np.clip(c[:, 1] - side[:, 1] / 2, 0, S - 2)
        y2, x2 = np.clip(c[:, 0] + side[:, 0] / 2, y1 + 1, S), np.clip(c[:, 1] + side[:, 1] / 2, x1 + 1, S)
        box[b, rows] = np.round(np.stack([y1, x1, y2, x2], 1)).astype(np.int32)
        cls[b, rows] = rng.integers(1, 81, n_real)
        cls[b, rows[:crowds]] *= -1
    if B > 1:
        box[1, np.nonzero(cls[1] > 0)[0][0]] = [S // 2, S // 4, S // 2, S // 2]   # zero-area instance: all anchors tie
    keys = rng.random((B, A), dtype=np.float32)
    keys[0, : A // 2] = np.round(keys[0, : A // 2] * 8) / 8               # heavy key ties on image 0
    want = orc.build_rpn_targets(an, cls, box, keys, R, SD)
    match, bbox, counts = _cuda_targets(dev, an, cls, box, keys, R, return_counts=True)
    assert np.array_equal(match, want["rpn_match"])
    assert np.array_equal(counts, want["counts"])
    check_bbox(bbox, want["rpn_bbox"])


@pytest.mark.gpu
def test_reference_signature_single_image(G, dev):
    import torch
    from maskrcnn_tf2_b200.targets import build_rpn_targets
    an, cls, box, keys, R = case(G, "small")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    real = cls[0] != 0                                                    # the loader passes only the real instances
    match, bbox = build_rpn_targets(t(an), t(cls[0][real]), t(box[0][real]), R, np.array(SD), rand_keys=t(keys[0]))
    assert match.shape == (an.shape[0],) and match.dtype == torch.int32 and bbox.shape == (R, 4)
    assert np.array_equal(match.cpu().numpy(), G["small_rpn_match"][0].astype(np.int32))
    check_bbox(bbox.cpu().numpy(), G["small_rpn_bbox"][0])
    match2, _ = build_rpn_targets(t(an), t(cls), t(box), R, SD)          # batched, keys from torch's CUDA generator
    m = match2.cpu().numpy()
    assert m.shape == cls.shape[:1] + (an.shape[0],) and ((m == 1).sum(1) <= R // 2).all() and ((m != 0).sum(1) <= R).all()
