import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maskrcnn_tf2_b200 import functional as F
rng = np.random.default_rng(0)
B, M = 2, 3000
c = rng.uniform(0.1, 0.9, (B, M, 2)); s = rng.uniform(0.02, 0.2, (B, M, 2))
boxes = np.concatenate([c - s / 2, c + s / 2], -1).astype(np.float32)
scores = rng.uniform(0, 1, (B, M)).astype(np.float32)
keep, count = F.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 500, 0.5)
torch.cuda.synchronize()
print("count", count.cpu().numpy())
