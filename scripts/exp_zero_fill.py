"""Experiment: streaming component of the deterministic ROIAlign backward (all ROIs out of range -> pure zero-fill)
against torch's memset of the same bytes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maskrcnn_tf2_b200 import functional as F
dev = torch.device("cuda:0")
B, T, C = 8, 200, 256
shapes = [(B, s, s, C) for s in (256, 128, 64, 32)]
def timed(fn, reps=20):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2] * 1e3
bufs = [torch.empty(s, device=dev) for s in shapes]
print("torch zero_ of the four maps: %.0f us" % timed(lambda: [b.zero_() for b in bufs]))
boxes = torch.tensor([2.0, 2.0, 2.5, 2.5], device=dev).repeat(B, T, 1).contiguous()
roi_map = torch.zeros((B, T), dtype=torch.int32, device=dev)
g = torch.randn((B, T, 7, 7, C), device=dev)
print("gather backward, no samples: %.0f us" % timed(lambda: F.roialign_backward(g, boxes, roi_map, shapes)))
print("atomic backward, no samples: %.0f us" % timed(lambda: F.roialign_backward(g, boxes, roi_map, shapes, deterministic=False)))
