"""Seeded stress sweep of the three kernels rebuilt in round 2 (NMS sweep, top-k, small-N detection NMS) against the CPU
oracle: random batch sizes, list lengths, thresholds, valid counts, score distributions (ties, floods, saturated, -inf).
    python scripts/stress_nms_topk.py [seconds] [seed]        (test infrastructure: imports oracle/)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import oracle
from conftest import random_boxes
from maskrcnn_tf2_b200 import functional as F

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20261019)
dev = torch.device("cuda:0")
oracle.build()
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
t0 = time.time()
n_nms = n_topk = 0
while time.time() - t0 < budget:
    # ---- NMS (generic entry: sort + sweep / lazy / small path depending on the size)
    B = int(rng.integers(1, 13))
    M = int(rng.choice([40, 150, 257, 900, 2048, 2049, 2500, 3100, 4096, 5000, 6000, 8192]))
    boxes = np.stack([random_boxes(rng, M, clusters=int(rng.integers(0, 80))) for _ in range(B)])
    scores = rng.uniform(0, 1, (B, M)).astype(np.float32)
    kind = int(rng.integers(0, 5))
    if kind == 1:
        scores = (np.round(scores * 32) / 32).astype(np.float32)
    elif kind == 2:
        scores[:, rng.integers(0, M, M // 3)] = -np.inf
    elif kind == 3:
        boxes[:, M // 2:] = boxes[:, :M - M // 2]          # exact duplicates
    valid = None if rng.uniform() < 0.5 else [int(v) for v in rng.integers(0, M + 1, B)]
    thr = float(rng.choice([0.1, 0.3, 0.5, 0.7, 0.9]))
    max_out = int(rng.choice([1, 10, 100, 1000, 2000, M]))
    keep, count = F.nms(T(boxes), T(scores), max_out, thr, None if valid is None else T(np.asarray(valid, np.int32)))
    keep, count = keep.cpu().numpy(), count.cpu().numpy()
    for b in range(B):
        n = M if valid is None else valid[b]
        ref = oracle.nms(boxes[b, :n], scores[b, :n], max_out, thr)
        assert count[b] == len(ref) and np.array_equal(keep[b, :len(ref)], ref), ("nms", B, M, kind, valid, thr, max_out, b)
    n_nms += 1
    # ---- top-k
    B = int(rng.integers(1, 10))
    A = int(rng.choice([1000, 4092, 16368, 65472, 261888]))
    K = int(min(A, rng.choice([1, 100, 1000, 6000, 8192])))
    kind = int(rng.integers(0, 5))
    s = rng.standard_normal((B, A, 2)).astype(np.float32)
    if kind == 1:
        s = 1.0 / (1.0 + np.exp(-4 * s))                                  # probabilities
    elif kind == 2:
        s = (np.round(s * 4) / 4).astype(np.float32)                      # heavy ties
    elif kind == 3:
        s[:, :, 1] = np.where(rng.uniform(size=(B, A)) < 0.7, 1.0, s[:, :, 1])   # saturated flood
    elif kind == 4:
        s[:, rng.integers(0, A, A // 4), 1] = -np.inf
    s = s.astype(np.float32)
    idx = F.topk(T(s), K, column=1).cpu().numpy()
    for b in range(B):
        assert np.array_equal(idx[b], oracle.topk(s[b, :, 1], K)), ("topk", B, A, K, kind, b)
    n_topk += 1
print(f"stress: {n_nms} NMS problems and {n_topk} top-k problems in {time.time() - t0:.0f} s, all bit-exact against the oracle")
