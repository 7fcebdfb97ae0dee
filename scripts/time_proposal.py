"""ProposalLayer device time vs batch size (cluster-size selection check).  python scripts/time_proposal.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maskrcnn_tf2_b200 import functional as F, synth
dev = torch.device("cuda:0")
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
a = synth.pyramid_anchors(1024)
base = [synth.rpn_outputs(np.random.default_rng(2000 + b), a, "clustered", 1024) for b in range(8)]
def timed(fn, n=20):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for x, y in ev:
        x.record(); fn(); y.record()
    torch.cuda.synchronize()
    return sorted(x.elapsed_time(y) for x, y in ev)[n // 2] * 1e3
for B in [int(v) for v in sys.argv[1:]] or [1, 2, 4, 8, 12, 16, 24, 32, 64]:
    pr = torch.from_numpy(np.stack([base[b % 8][0] for b in range(B)])).to(dev)
    bb = torch.from_numpy(np.stack([base[b % 8][1] for b in range(B)])).to(dev)
    an = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, (B,) + a.shape))).to(dev)
    us = timed(lambda: F.proposal_forward(pr, bb, an, 6000, 1000, SD, 0.7))
    print(f"B={B:3d}: ProposalLayer {us:7.1f} us  ({B / us * 1e6:8.0f} images/s)", flush=True)
