"""build_rpn_targets: device time of the batched CUDA path vs the CPU restatement of the reference's numpy routine.
python scripts/time_rpn_targets.py [B ...]   (COCO shape: 1024^2, A = 261888, G = 100 padded / 20 real, R = 256)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from maskrcnn_tf2_b200 import functional as F, synth
dev = torch.device("cuda:0")
SD = (0.1, 0.1, 0.2, 0.2)
S, G, R, n_real = 1024, 100, 256, 20
an = synth.pyramid_anchors_px(S)
A = an.shape[0]
def batch(B):
    cls, box = np.zeros((B, G), np.int32), np.zeros((B, G, 4), np.int32)
    for b in range(B):
        rng = np.random.default_rng(3000 + b)
        side, c = rng.uniform(32, 512, (n_real, 2)), rng.uniform(0, S, (n_real, 2))
        y1, x1 = np.clip(c[:, 0] - side[:, 0] / 2, 0, S - 2), np.clip(c[:, 1] - side[:, 1] / 2, 0, S - 2)
        y2, x2 = np.clip(c[:, 0] + side[:, 0] / 2, y1 + 2, S), np.clip(c[:, 1] + side[:, 1] / 2, x1 + 2, S)
        box[b, :n_real] = np.round(np.stack([y1, x1, y2, x2], 1))
        cls[b, :n_real] = rng.integers(1, 81, n_real)
    return cls, box, np.random.default_rng(B).random((B, A), dtype=np.float32)
def timed(fn, n=20):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for x, y in ev:
        x.record(); fn(); y.record()
    torch.cuda.synchronize()
    return sorted(x.elapsed_time(y) for x, y in ev)[n // 2] * 1e3
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
tan = t(an)
for B in [int(v) for v in sys.argv[1:]] or [1, 8, 32]:
    cls, box, keys = batch(B)
    tc, tb, tk = t(cls), t(box), t(keys)
    us = timed(lambda: F.rpn_targets_forward(tan, tc, tb, tk, R, SD, return_f32=True))
    t0 = time.perf_counter(); want = oracle.build_rpn_targets(an, cls, box, keys, R, SD); cpu = time.perf_counter() - t0
    got = F.rpn_targets_forward(tan, tc, tb, tk, R, SD)[0].cpu().numpy()
    print(f"B={B:3d}: CUDA {us:8.1f} us ({B / us * 1e6:9.0f} images/s)   CPU restatement ({oracle.max_threads()} threads) "
          f"{cpu * 1e3:8.1f} ms ({B / cpu:7.1f} images/s)   rpn_match equal: {np.array_equal(got, want['rpn_match'])}", flush=True)
