"""Phase timeline of topk_cluster_kernel, thread 0 of CTA 0 (debug build `make -C maskrcnn_tf2_b200/csrc prof`)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maskrcnn_tf2_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ.get("MRCNN_PROF_LIB", "libmrcnn_roi_b200_prof.so"))
from maskrcnn_tf2_b200 import functional as F, synth
L = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
a = synth.pyramid_anchors(1024)
pr, bb = zip(*[synth.rpn_outputs(np.random.default_rng(2000 + b), a, "clustered", 1024) for b in range(B)])
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
probs, bbox, anch = t(np.stack(pr)), t(np.stack(bb)), t(np.broadcast_to(a, (B,) + a.shape))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    out = F.proposal_forward(probs, bbox, anch, 6000, 1000, [0.1, 0.1, 0.2, 0.2], 0.7)
torch.cuda.synchronize()
tl = (ctypes.c_longlong * 32)()
L.mrcnn_debug_topk_timeline.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
L.mrcnn_debug_topk_timeline(tl)
tl = np.array(list(tl), dtype=np.int64)
names = {0: "entry", 1: "dependency wait over", 26: "compaction done", 27: "(candidate list complete)", 28: "cluster barrier (lists)",
         29: "all lists gathered", 30: "ranked by counting", 31: "epilogue done"}
for lv in range(6):
    names[2 + 4 * lv] = f"level {lv}: histogram pass done"
    names[3 + 4 * lv] = f"level {lv}: cluster barrier"
    names[4 + 4 * lv] = f"level {lv}: peers' histograms summed"
    names[5 + 4 * lv] = f"level {lv}: digit resolved"
prev = tl[0]
for i in sorted(names):
    if tl[i] >= tl[0] and tl[i] > 0:
        print(f"  {names[i]:40s} {tl[i] - tl[0]:8d}  (+{tl[i] - prev})")
        prev = tl[i]
