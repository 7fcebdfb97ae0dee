"""Per-tile timeline of nms_sweep_kernel, CTA 0 (debug build `make -C maskrcnn_tf2_b200/csrc prof`)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maskrcnn_tf2_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ.get("MRCNN_PROF_LIB", "libmrcnn_roi_b200_prof.so"))
from maskrcnn_tf2_b200 import functional as F, synth
L = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
regime = sys.argv[2] if len(sys.argv) > 2 else "clustered"
a = synth.pyramid_anchors(1024)
pr, bb = zip(*[synth.rpn_outputs(np.random.default_rng(2000 + b), a, regime, 1024) for b in range(B)])
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
probs, bbox, anch = t(np.stack(pr)), t(np.stack(bb)), t(np.broadcast_to(a, (B,) + a.shape))
for _ in range(3):
    out = F.proposal_forward(probs, bbox, anch, 6000, 1000, [0.1, 0.1, 0.2, 0.2], 0.7, debug=True)
torch.cuda.synchronize()
tl = (ctypes.c_longlong * (128 * 8))()
L.mrcnn_debug_nms_sweep_timeline.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
L.mrcnn_debug_nms_sweep_timeline(tl)
tl = np.array(list(tl), dtype=np.int64).reshape(128, 8)
t0 = tl[0, 0]
print(f"B={B} {regime}: timeline of CTA 0 (cycles since resolve(0) started); columns: resolver start | data arrived | kept known | "
      "released || far warp 0 of the owning CTA sent (tiles of CTA 0 only) || row warp 0 sent a job of the tile (CTA 0 only)")
last = max(i for i in range(120) if tl[i, 3] > 0)
for i in range(0, last + 1):
    r = tl[i] - t0
    print(f"  tile {i}: {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} || {r[5] if tl[i,5] else -1:7d} || {r[6] if tl[i,6] else -1:7d}")
g = tl[120] - t0
print(f"  kernel entry {g[0]}, after the dependency wait {g[1]}, prologue jobs done {g[2]}, cluster barrier passed {g[3]}, sweep over {g[4]}, "
      f"drained {g[5]}, outputs written {g[6]}")
e = tl[121] - t0
lanes = tl[64:96, 7] - t0
print(f"  CTA 0: tail far warp 0 left its loop at {e[0]}, bulk far warp 0 at {e[1]}, row warp 0 at {e[2]}; the drain's barrier waits "
      f"ended per lane between {lanes.min()} and {lanes.max()} (lane {int(lanes.argmax())})")
print(f"  tiles {last + 1}: {(tl[last, 3] - t0) / (last + 1):.0f} cycles per tile")

rt = (ctypes.c_longlong * (128 * 4))()
L.mrcnn_debug_nms_sweep_rows.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
L.mrcnn_debug_nms_sweep_rows(rt)
rt = np.array(list(rt), dtype=np.int64).reshape(128, 4)
print("row warp 0 of CTA 0, per job: gate passed | boxes staged | rows done | sent  (cycles since resolve(0) started)")
for i in range(128):
    if rt[i, 3] > 0:
        r = rt[i] - t0
        print(f"  tile {i}: {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d}")
