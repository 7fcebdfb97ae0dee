#!/usr/bin/env python
"""ROI-stage benchmark (BASELINE.json: "ROI-stage images/sec (Proposal+ROIAlign+Detection)").

    python bench.py --gpus N --steps K --warmup W [--config C]   our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference ...                          the reference arm: the CPU restatement of the
                                                                  reference's TF path (oracle/), all host threads

--config selects the BASELINE.json configuration (default 2 = configs[1], the one the metric is quoted on):
    1  configs[0]  balloon ResNet-50 FPN inference, batch 1, 1024^2, 2 classes
    2  configs[1]  COCO-shape ROI stage, batch 8 per GPU, 1024^2, A=261888, 6000 -> 1000 RoIs, 81 classes   (weak scaling)
    3  configs[2]  COCO-shape training step: DetectionTargetLayer (2000 proposals -> T=200) + PyramidROIAlign forward AND
                   backward at 7x7 and 14x14, batch 8 per GPU
    4  configs[3]  small feature maps: batch 32 per GPU, 512^2 (--img-size 256 for the other reading)
    5  configs[4]  batch 64 in total, sharded over the N GPUs (64 / 32 / 16 / 8 images each)              (strong scaling)

A step of an inference configuration = one pass of the inference ROI stage over one batch on each GPU:
    ProposalLayer -> PyramidROIAlign 7x7 (N=1000) -> DetectionLayer -> PyramidROIAlign 14x14 (N=100)
(the classifier / mask heads between them stay in TensorFlow and are represented by their synthetic outputs).

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "roi_stage_images_per_sec"
UNIT = "images/s"

CONFIGS = {
    1: dict(kind="inference", batch=1, img=1024, nc=2, scaling="weak",
            workload="configs[0]: balloon-config ResNet-50 FPN inference ROI stage, batch 1, 1024x1024, A=261888, 6000 "
                     "pre-NMS / 1000 post-NMS RoIs, 2-class DetectionLayer, clustered RPN regime"),
    2: dict(kind="inference", batch=8, img=1024, nc=81, scaling="weak",
            workload="configs[1]: COCO-shape ResNet-101 FPN ROI stage, batch 8/GPU, 1024x1024, A=261888, 6000 pre-NMS / "
                     "1000 post-NMS RoIs, 81-class DetectionLayer, clustered RPN regime"),
    3: dict(kind="training", batch=8, img=1024, nc=81, scaling="weak",
            workload="configs[2]: COCO-shape training step, batch 8/GPU, 1024x1024: DetectionTargetLayer (2000 proposals, "
                     "100 GT slots / 20 real, full 1024^2 masks -> T=200) + PyramidROIAlign forward and backward at 7x7 "
                     "and 14x14 (deterministic gradient)"),
    4: dict(kind="inference", batch=32, img=512, nc=81, scaling="weak",
            workload="configs[3]: small-feature-map ROI stage (MobileNetV2 / EfficientNet-B0 class backbones, FPN width "
                     "256), batch 32/GPU, 512x512, A=65472, 81 classes, clustered RPN regime"),
    5: dict(kind="inference", total_batch=64, img=1024, nc=81, scaling="strong",
            workload="configs[4]: COCO-shape ROI stage, batch 64 in total sharded over the GPUs (64/N images each, "
                     "contiguous, distinct images), 1024x1024, 81 classes, clustered RPN regime"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default: the configuration's)")
    ap.add_argument("--total-batch", type=int, default=None, help="config 5: images per step over all GPUs (default 64)")
    ap.add_argument("--img-size", type=int, default=None)
    ap.add_argument("--num-classes", type=int, default=None)
    ap.add_argument("--regime", default="clustered", choices=["clustered", "sparse", "iid"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true",
                    help="`value` from eager layer calls instead of the captured CUDA graph of the same calls")
    ap.add_argument("--e2e-full-copy", action="store_true",
                    help="e2e leg: bulk-copy the feature maps to HBM every step instead of demand-fetching the sampled pixels")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bound on the CPU baseline sample")
    ap.add_argument("--no-pipelined", action="store_true", help="skip the extra two-stream throughput measurement")
    ap.add_argument("--first-image", type=int, default=0, help="index of the first synthetic image")
    ap.add_argument("--distinct-images", action="store_true",
                    help="rank r processes images first+r*batch.. instead of the same batch as every other rank")
    return ap.parse_args()


def resolve(args, world):
    """The workload this run measures: identical for both arms (the `config` object of the JSON line)."""
    c = dict(CONFIGS[args.config])
    if c["scaling"] == "strong":
        total = args.total_batch or c["total_batch"]
        if total % world:
            raise SystemExit(f"--config {args.config}: {total} images do not split evenly over {world} GPUs")
        c["batch"] = total // world
        c["total_batch"] = total
    if args.batch:
        c["batch"] = args.batch
    if args.img_size:
        c["img"] = args.img_size
    if args.num_classes:
        c["nc"] = args.num_classes
    c["distinct"] = args.distinct_images or c["scaling"] == "strong"
    return c


def config_object(args, c, world):
    """Same keys and values from both arms (the driver compares the two lines' `config`)."""
    return {"workload": c["workload"], "config_id": args.config, "kind": c["kind"], "batch_per_gpu": c["batch"],
            "global_batch": c["batch"] * world, "img_size": c["img"], "num_classes": c["nc"], "regime": args.regime,
            "scaling": c["scaling"], "first_image": args.first_image,
            "images": "distinct images per rank" if c["distinct"] else "the same synthetic batch on every rank",
            "l2": "inputs larger than L2: the feature maps of one step (89 MB/image at 1024^2, 22 MB at 512^2) exceed "
                  "the 126 MB L2 at every batch size used"}


def algorithmic_bytes(img_size, n_rois, ph, pw, C=256):
    """SURVEY.md 8(d) / BASELINE.md section 3: ROIAlign forward bytes per image (closed form, upper bound)."""
    maps = sum((img_size // s) ** 2 * C * 4 for s in (4, 8, 16, 32))
    out = n_rois * ph * pw * C * 4
    return out + min(4 * out, maps) + 16 * n_rois


def backward_algorithmic_bytes(img_size, n_rois, ph, pw, C=256):
    """ROIAlign backward bytes per image = gradient read once + every gradient-map pixel written once."""
    maps = sum((img_size // s) ** 2 * C * 4 for s in (4, 8, 16, 32))
    return n_rois * ph * pw * C * 4 + maps + 16 * n_rois


def touched_map_bytes(boxes, roi_map, img_size, ph, pw, C=256):
    """Compulsory feature-map reads of one PyramidROIAlign launch: every (image, map, pixel) some ROI samples,
    counted once (C*4 bytes each).  Restates the tap arithmetic of TF crop_and_resize in float32 numpy."""
    import numpy as np
    B, N, _ = boxes.shape
    sizes = [img_size // s for s in (4, 8, 16, 32)]
    total = 0
    for b in range(B):
        for m in range(4):
            sel = boxes[b][roi_map[b] == m]
            if sel.shape[0] == 0:
                continue
            H = W = sizes[m]
            f32 = np.float32
            ys = sel[:, 0:1] * f32(H - 1) + np.arange(ph, dtype=f32)[None, :] * ((sel[:, 2:3] - sel[:, 0:1]) * f32(H - 1) / f32(ph - 1))
            xs = sel[:, 1:2] * f32(W - 1) + np.arange(pw, dtype=f32)[None, :] * ((sel[:, 3:4] - sel[:, 1:2]) * f32(W - 1) / f32(pw - 1))
            vy = (ys >= 0) & (ys <= H - 1)
            vx = (xs >= 0) & (xs <= W - 1)
            y0, y1 = np.floor(ys).astype(np.int64), np.ceil(ys).astype(np.int64)
            x0, x1 = np.floor(xs).astype(np.int64), np.ceil(xs).astype(np.int64)
            seen = np.zeros((H, W), dtype=bool)
            for yy, xx in ((y0, x0), (y0, x1), (y1, x0), (y1, x1)):
                valid = vy[:, :, None] & vx[:, None, :]
                Y = np.broadcast_to(np.clip(yy, 0, H - 1)[:, :, None], valid.shape)[valid]
                X = np.broadcast_to(np.clip(xx, 0, W - 1)[:, None, :], valid.shape)[valid]
                seen[Y, X] = True
            total += int(seen.sum()) * C * 4
    return total


def stage_bytes(img_size, A, K=6000, P=1000, NC=81, D=100):
    """SURVEY.md 8(d) closed form for the inference stage, per image (its ROIAlign terms assume every pixel touched)."""
    prop = A * 8 + 2 * K * 16 + P * 16
    det = P * 16 + P * NC * 4 + P * NC * 16 + D * 24
    return prop + det + algorithmic_bytes(img_size, P, 7, 7) + algorithmic_bytes(img_size, D, 14, 14)


def source_hash():
    """SHA-256 over the CUDA sources the library is built from: the key of profiles/roofline_traffic.json."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "maskrcnn_tf2_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh")):
            h.update(name.encode())
            h.update(open(os.path.join(csrc, name), "rb").read())
    return h.hexdigest()


def measured_traffic(config_id, regime, key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch / per step from the committed ncu capture of THIS build
    (profiles/roofline_traffic.json, written by scripts/ncu_traffic.py from an `ncu --set full` run); None when the
    capture was made from other sources (stale) or does not hold this configuration."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if rec.get("source_sha256") != source_hash():
            return None
        return rec["configs"][f"{config_id}:{regime}"].get(key)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (NVML; nvidia-smi as fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _sample_nvml(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
        for bit, name in names.items():
            if bits & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        import subprocess
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.check_output(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                       "--format=csv,noheader,nounits"], timeout=5).decode().strip().split(",")
        self.samples.append(int(out[0]))
        self.max_mhz = int(out[1])
        for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:]):
            if v.strip().lower() == "active":
                self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nv is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def merge_clocks(a, b):
    """Clock record over two timed regions (the graph replay that gives `value` and the instrumented pass)."""
    if a is None:
        return b
    s = [v for v in (a["sm_mhz"], b["sm_mhz"]) if v is not None]
    return {"sm_mhz": min(s) if s else None, "sm_max_mhz": a["sm_max_mhz"] or b["sm_max_mhz"],
            "reasons": sorted(set(a["reasons"]) | set(b["reasons"])), "samples": a["samples"] + b["samples"]}


def bind_to_gpu_cpus(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index` (same NUMA node / PCIe root), so the
    pinned host buffers of the end-to-end loop are allocated next to the GPU that reads them.  Returns the previous
    affinity (restored around the CPU baseline, which uses every core) or None."""
    try:
        import pynvml
        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        handle = None
        try:                                   # CUDA and NVML orderings can differ: match by UUID when torch has it
            import torch
            handle = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(index).uuid))
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return before
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------
TRAIN_T, TRAIN_P = 200, 2000
SD4 = (0.1, 0.1, 0.2, 0.2)


def make_inputs(args, c, first):
    """Synthetic host inputs of one rank (numpy).  Training adds the data loader's GT tensors, the shuffle keys and the
    head gradients that flow back into the two PyramidROIAlign layers."""
    import numpy as np
    from maskrcnn_tf2_b200 import synth
    B, S, NC = c["batch"], c["img"], c["nc"]
    x = synth.inference_batch(args.config, B, img_size=S, num_classes=NC, regime=args.regime, first_image=first)
    if c["kind"] == "training":
        g = synth.training_targets_batch(args.config, B, img_size=S, first_image=first)
        x.update(g)
        x["rand_keys"] = np.random.default_rng(9 + first).integers(0, 2 ** 32, (B, TRAIN_P), dtype=np.uint64).astype(np.uint32)
        rng = np.random.default_rng(10 + first)
        x["grad7"] = rng.standard_normal((B, TRAIN_T, 7, 7, 256), dtype=np.float32)
        x["grad14"] = rng.standard_normal((B, TRAIN_T, 14, 14, 256), dtype=np.float32)
    return x


def cpu_stage(oracle, x, cfg):
    """The reference's CPU path for one batch (oracle = restatement of the TF kernels + layer control flow)."""
    import numpy as np
    S = float(cfg["img_size"])
    r = oracle.proposal_layer(x["rpn_probs"], x["rpn_bbox"], x["anchors"], cfg["pre_nms_limit"],
                              cfg["post_nms_rois_inference"], cfg["rpn_bbox_std_dev"], cfg["rpn_nms_threshold"])
    oracle.pyramid_roi_align(r["proposals"], S, S, x["feature_maps"], (7, 7))
    d = oracle.detection_layer(r["proposals"], x["mrcnn_class"], x["mrcnn_bbox"], x["image_meta"], cfg["bbox_std_dev"],
                               cfg["detection_min_confidence"], cfg["detection_max_instances"],
                               cfg["detection_nms_threshold"])
    oracle.pyramid_roi_align(np.ascontiguousarray(d["detections"][..., :4]), S, S, x["feature_maps"], (14, 14))
    return d["detections"]


def cpu_training_stage(oracle, x, cfg, proposals):
    S = float(cfg["img_size"])
    shapes = [f.shape for f in x["feature_maps"]]
    t = oracle.detection_target_layer(proposals, x["gt_class_ids"], x["gt_boxes"], x["gt_masks"], x["rand_keys"],
                                      TRAIN_T, cfg["roi_positive_ratio"], cfg["bbox_std_dev"], (28, 28))
    for pool, g in (((7, 7), x["grad7"]), ((14, 14), x["grad14"])):
        oracle.pyramid_roi_align(t["rois"], S, S, x["feature_maps"], pool)
        oracle.pyramid_roi_align_grad(g, t["rois"], S, S, shapes)
    return t["rois"]


def time_cpu(x, cfg, c, max_seconds, steps=None, warmup=1):
    """Times the oracle on the host cores.  Returns (images/s, seconds per batch, reps, threads)."""
    import oracle
    threads = os.cpu_count() or 1
    oracle.set_num_threads(threads)
    if c["kind"] == "training":
        props = oracle.proposal_layer(x["rpn_probs"], x["rpn_bbox"], x["anchors"], cfg["pre_nms_limit"], TRAIN_P,
                                      cfg["rpn_bbox_std_dev"], cfg["rpn_nms_threshold"])["proposals"]
        run = lambda: cpu_training_stage(oracle, x, cfg, props)
    else:
        run = lambda: cpu_stage(oracle, x, cfg)
    for _ in range(warmup):
        run()
    times = []
    t_begin = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() - t_begin > max_seconds or len(times) >= 30:
            break
    times.sort()
    med = times[len(times) // 2]
    return c["batch"] / med, med, len(times), threads


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  TensorFlow (tensorflow==2.2-2.5,
    the reference's only implementation) cannot be installed here, so this is the oracle port, all host threads.
    Rank 0 alone runs it, on the shard rank 0 of our arm processes."""
    if rank != 0:
        return
    from maskrcnn_tf2_b200 import make_config
    c = resolve(args, world)
    cfg = make_config(img_size=c["img"], num_classes=c["nc"], batch_size=c["batch"])
    x = make_inputs(args, c, args.first_image)
    ips, sec, reps, threads = time_cpu(x, cfg, c, args.cpu_seconds, steps=args.steps, warmup=max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": reps,
        "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": c["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_object(args, c, world),
        "note": "CPU restatement of the reference's TF path (TensorFlow unavailable offline); one process, OpenMP over "
                "images/ROIs, on the batch one GPU of our arm processes",
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{reps} x one per-GPU batch of {c['batch']} images, median"},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
_JSON_OUT = None


def claim_stdout():
    """stdout must carry exactly one JSON line.  Libraries loaded later (NCCL prints its version banner with printf)
    write to file descriptor 1 behind Python's back, so keep a private duplicate of the real stdout for the JSON line
    and point descriptor 1 at stderr for everybody else."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    args = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from maskrcnn_tf2_b200 import functional as F
    from maskrcnn_tf2_b200 import make_config
    from maskrcnn_tf2_b200.graphs import CapturedStage
    from maskrcnn_tf2_b200.layers import (DetectedBoxesExtraction, DetectionLayer, DetectionTargetLayer, ProposalLayer,
                                          PyramidROIAlign)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ROI-stage kernels have no CPU path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = bind_to_gpu_cpus(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    c = resolve(args, world)
    B, S, NC = c["batch"], c["img"], c["nc"]
    training = c["kind"] == "training"
    cfg = make_config(img_size=S, num_classes=NC, batch_size=B)
    # weak scaling: every GPU processes B images per step; by default every rank gets the SAME synthetic batch, so that
    # per-GPU work is exactly fixed as N grows (--distinct-images gives rank r its own).  Strong scaling (config 5): the
    # global batch is split contiguously with sharding.shard_range, every rank has its own images.
    if c["scaling"] == "strong":
        from maskrcnn_tf2_b200.sharding import shard_range
        lo, hi = shard_range(c["total_batch"], rank, world)
        assert hi - lo == B
        first = args.first_image + lo
    else:
        first = args.first_image + (rank * B if c["distinct"] else 0)
    x = make_inputs(args, c, first)
    A = x["anchors"].shape[1]

    def pin(a):
        return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).pin_memory()
    host = {k: pin(v) for k, v in x.items() if k != "feature_maps"}
    host_maps = [pin(f) for f in x["feature_maps"]]
    d = {k: v.to(dev) for k, v in host.items()}
    d_maps = [f.to(dev) for f in host_maps]
    anchors = d["anchors"]           # model constant (AnchorsLayer's non-trainable variable): always resident
    P_ = cfg["post_nms_rois_inference"]

    align7 = PyramidROIAlign([cfg["pool_size"]] * 2, name="roi_align_classifier")
    align14 = PyramidROIAlign([cfg["mask_pool_size"]] * 2, name="roi_align_mask")
    timed_steps = args.steps
    ev_roof = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(timed_steps)]
    map_shapes = [tuple(f.shape) for f in d_maps]

    if not training:
        proposal = ProposalLayer(P_, cfg)
        detect = DetectionLayer(P_, cfg["detection_min_confidence"], cfg["detection_max_instances"],
                                cfg["detection_nms_threshold"], cfg["bbox_std_dev"], B, B)
        boxes_of = DetectedBoxesExtraction(cfg)
        # our kernels per step: proposal (topk_cluster + nms_lazy), align7 (prep + fwd), detection (refine + nms_lazy
        # with fused ordering, which also writes detections[..., :4] for the mask branch), align14 (prep + fwd)
        KERNELS_PER_STEP = 2 + 2 + 2 + 2

        def stage(t, maps, step=None, host_stage=None):
            rois = proposal([t["rpn_probs"], t["rpn_bbox"], anchors])
            if step is not None:
                ev_roof[step][0].record()
            pooled = align7([rois, t["image_meta"]] + maps, host_stage=host_stage, new_maps=True)
            if step is not None:
                ev_roof[step][1].record()
            det = detect([rois, t["mrcnn_class"], t["mrcnn_bbox"], t["image_meta"]])
            mask_pooled = align14([boxes_of(det), t["image_meta"]] + maps, host_stage=host_stage, new_maps=False)
            return rois, pooled, det, mask_pooled
    else:
        tcfg = dict(cfg, train_rois_per_image=TRAIN_T)
        targets = DetectionTargetLayer(tcfg)
        proposals = F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], anchors, cfg["pre_nms_limit"], TRAIN_P,
                                       cfg["rpn_bbox_std_dev"], cfg["rpn_nms_threshold"])
        d["proposals"] = proposals
        host["proposals"] = proposals.cpu().pin_memory()
        # dt_select + dt_mask, 2 x (prep + fwd), 2 x deterministic backward (const + count + alloc + fill + gather +
        # fallback scatter; the memset node in front of them is not counted)
        KERNELS_PER_STEP = 2 + 2 * 2 + 2 * 6

        def stage(t, maps, step=None, host_stage=None):
            rois, cls, deltas, masks = targets([t["proposals"], t["gt_class_ids"], t["gt_boxes"], t["gt_masks"]],
                                               rand_keys=t["rand_keys"])
            out7, map7 = F.roialign_forward(rois, t["image_meta"], maps, (7, 7))
            g7 = F.roialign_backward(t["grad7"], rois, map7, map_shapes, deterministic=True)
            out14, map14 = F.roialign_forward(rois, t["image_meta"], maps, (14, 14))
            if step is not None:
                ev_roof[step][0].record()
            g14 = F.roialign_backward(t["grad14"], rois, map14, map_shapes, deterministic=True)
            if step is not None:
                ev_roof[step][1].record()
            return rois, out7, g7, out14, g14, cls, deltas, masks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: inputs in HBM ----
    # (1) instrumented pass: eager layer calls, CUDA events around the dominant kernel's layer call in every step
    for _ in range(max(args.warmup, 3)):
        outs = stage(d, d_maps)          # same allocation pattern as the timed loop (previous outputs still alive)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(timed_steps):
        outs = stage(d, d_maps, i)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_eager = e0.elapsed_time(e1)
    ms_roof = sorted(a.elapsed_time(b) for a, b in ev_roof)
    ms_roof_avg = sum(ms_roof) / len(ms_roof)

    # (2) `value`: the same layer calls captured once into a CUDA graph (graphs.CapturedStage) and replayed K times --
    # one host call per step, the programmatic-dependent-launch edges between the kernels kept inside the graph
    ms_total, graph_note = ms_eager, "eager layer calls"
    if not args.no_graph:
        captured = CapturedStage(lambda: stage(d, d_maps), warmup=2, device=dev)
        for _ in range(max(args.warmup, 3)):
            captured.replay()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(timed_steps):
            captured.replay()
        g1.record()
        barrier()
        clocks = merge_clocks(sampler.stop(), clocks)
        ms_total = g0.elapsed_time(g1)
        graph_note = "CUDA graph of the layer calls (graphs.CapturedStage), one replay per step"
        gouts = captured.outputs
        for a_, b_ in zip(outs[:4], gouts[:4]):        # the replayed graph computes what the eager calls compute
            ta, tb = (a_, b_) if isinstance(a_, torch.Tensor) else (a_[0], b_[0])
            assert torch.equal(ta, tb), "graph replay and eager layer calls disagree"

    # ---- extra: the same K steps issued alternately on two CUDA streams (batch i+1's latency-bound ProposalLayer
    # overlaps batch i's HBM-bound ROIAlign); reported beside `value`, which stays the single-stream number ----
    pipelined = None
    if not args.no_pipelined and not training:
        streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        keep_alive = [None, None]

        def run_pipelined(n):
            for i in range(n):
                # launchers own no state: each stream gets its own workspace set
                with F.workspace_namespace(1 + (i & 1)), torch.cuda.stream(streams[i & 1]):
                    keep_alive[i & 1] = stage(d, d_maps)
        for s_ in streams:
            s_.wait_stream(torch.cuda.current_stream())
        run_pipelined(max(args.warmup, 4))
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for s_ in streams:
            s_.wait_event(p0)
        run_pipelined(timed_steps)
        for s_ in streams:
            torch.cuda.current_stream().wait_stream(s_)
        p1.record()
        barrier()
        ms_pipe = p0.elapsed_time(p1)
    else:
        args.no_pipelined = True

    # ---- end to end: pinned host inputs -> H2D every step, result -> host every step ----
    e2e = None
    t_e2e = 0.0
    if not args.no_e2e and not training:
        # Host buffers in, detections out, every step.  The small inputs (RPN outputs, head outputs, image_meta: 63 MB)
        # are copied on a copy stream, the copy of step i+1 overlapping the kernels of step i.  The feature maps
        # (713 MB per step) stay in pinned host memory: each PyramidROIAlign call first fetches exactly the map pixels
        # its ROIs sample, once each, straight out of the host maps (F.HostMapStage / mrcnn_roialign_fetch_hostmaps),
        # so the PCIe traffic of a step is the sampled pixels, not the maps.  --e2e-full-copy restores the bulk copy.
        full_copy = args.e2e_full_copy
        # rpn_bbox [B,A,4] and mrcnn_bbox [B,N,NC,4] are read sparsely by index (the K top-k winners' rows; one class
        # row per ROI): they too stay in pinned host memory and only those rows cross the bus
        sparse = () if full_copy else ("rpn_bbox", "mrcnn_bbox")
        small = {k: v for k, v in host.items() if k != "anchors" and k not in sparse}
        sets = [{k: torch.empty_like(v, device=dev) for k, v in small.items()} for _ in range(2)]
        for s_ in sets:
            for k in sparse:
                s_[k] = host[k]
        sparse_bytes = 0 if full_copy else B * (min(cfg["pre_nms_limit"], A) + P_) * 16
        map_sets = [[torch.empty_like(f, device=dev) for f in host_maps] for _ in range(2)] if full_copy else None
        stages = None if full_copy else [F.HostMapStage(host_maps, dev) for _ in range(2)]
        det_host = [torch.empty((B, cfg["detection_max_instances"], 6), dtype=torch.float32).pin_memory()
                    for _ in range(2)]
        small_bytes = sum(v.numel() * v.element_size() for v in small.values())
        map_bytes = sum(f.numel() * f.element_size() for f in host_maps)
        pixel_bytes = host_maps[0].shape[3] * 4
        d2h = det_host[0].numel() * det_host[0].element_size()
        copy_stream = torch.cuda.Stream(device=dev)
        # two batches in flight: step i+1's ProposalLayer + pixel fetch (the bus) next to step i's ROIAlign / DetectionLayer
        compute_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        # A/B switch BENCH_E2E_STREAMS=2: two batches in flight (step i+1's ProposalLayer + pixel fetch next to step i's
        # ROIAlign / DetectionLayer).  Measured: +1 % at one GPU (1062 against 1053 images/s), +1 % at two (2129 against
        # 2107); with the fetch kernel at 4 CTAs per SM +2.5 % at one GPU but unstable at two (1400-1630).  Default: one.
        if os.environ.get("BENCH_E2E_STREAMS", "1") != "2":
            compute_streams[1] = compute_streams[0]
        compute_stream = compute_streams[0]
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        alive = [None, None]

        def issue_copy(i):
            dd = sets[i & 1]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i & 1])          # the buffers' previous step has been computed
                for k in small:
                    dd[k].copy_(small[k], non_blocking=True)
                if full_copy:
                    for a_, b_ in zip(map_sets[i & 1], host_maps):
                        a_.copy_(b_, non_blocking=True)
                copied[i & 1].record(copy_stream)

        def issue_compute(i):
            dd = sets[i & 1]
            cs_ = compute_streams[i & 1]
            with torch.cuda.stream(cs_), F.workspace_namespace(3 + (i & 1)):
                cs_.wait_event(copied[i & 1])
                if full_copy:
                    o = stage(dd, map_sets[i & 1])
                else:
                    o = stage(dd, host_maps, host_stage=stages[i & 1])
                alive[i & 1] = o
                consumed[i & 1].record(cs_)
                det_host[i & 1].copy_(o[2], non_blocking=True)
                done[i & 1].record(cs_)

        def run_e2e(n):
            issue_copy(0)
            for i in range(n):
                if i + 1 < n:
                    issue_copy(i + 1)
                issue_compute(i)
                if i >= 1:
                    done[(i - 1) & 1].synchronize()              # the caller reads step i-1's detections
            done[(n - 1) & 1].synchronize()

        for ev_, cs_ in zip(consumed, compute_streams):
            ev_.record(cs_)
        e2e_steps = max(3, min(args.steps, 30))
        run_e2e(3)
        barrier()
        fetched0 = 0 if full_copy else sum(s_.fetched_pixels() for s_ in stages)
        t0 = time.perf_counter()
        run_e2e(e2e_steps)
        barrier()
        t_e2e = time.perf_counter() - t0
        if full_copy:
            h2d = small_bytes + map_bytes
        else:   # bytes that actually crossed the bus: counted on the device by the fetch kernel
            h2d = small_bytes + sparse_bytes + \
                (sum(s_.fetched_pixels() for s_ in stages) - fetched0) * pixel_bytes / e2e_steps
        e2e_note = ("pinned host inputs copied to HBM and detections read back every step (copy of step i+1 overlaps the "
                    "kernels of step i on a second stream); anchors stay resident" if full_copy else
                    "pinned host inputs in, detections read back every step; RPN/head outputs copied to HBM (copy of step "
                    "i+1 overlaps step i), feature maps left in pinned host memory and only the pixels the ROIs sample "
                    "fetched over the bus (h2d_bytes_per_step counts them on the device; a full copy of all inputs would "
                    "be %d bytes); rpn_bbox / mrcnn_bbox are read in place, only the rows used; anchors stay resident"
                    % (small_bytes + map_bytes + sum(host[k].numel() * 4 for k in sparse)))
        e2e_maps = "full copy" if full_copy else "demand-fetched"
    elif not args.no_e2e:
        # training: what comes from the host every step is what the reference's data loader produces
        # (training.py:71-74): gt_class_ids, gt_boxes, gt_masks (full 1024^2 masks: 105 MB per image as tf.bool) and the
        # shuffle keys; proposals, feature maps and the head gradients are produced on the device by the network
        # around the path.  The target ROIs are read back every step.
        loader = ("gt_class_ids", "gt_boxes", "gt_masks", "rand_keys")
        sets = [dict(d, **{k: torch.empty_like(host[k], device=dev) for k in loader}) for _ in range(2)]
        rois_host = [torch.empty((B, TRAIN_T, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
        h2d = sum(host[k].numel() * host[k].element_size() for k in loader)
        d2h = rois_host[0].numel() * 4
        copy_stream, compute_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        alive = [None, None]

        def run_e2e(n):
            for i in range(n):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i & 1])
                    for k in loader:
                        sets[i & 1][k].copy_(host[k], non_blocking=True)
                    copied[i & 1].record(copy_stream)
                with torch.cuda.stream(compute_stream), F.workspace_namespace(3):
                    compute_stream.wait_event(copied[i & 1])
                    alive[i & 1] = stage(sets[i & 1], d_maps)
                    consumed[i & 1].record(compute_stream)
                    rois_host[i & 1].copy_(alive[i & 1][0], non_blocking=True)
                    done[i & 1].record(compute_stream)
                if i >= 1:
                    done[(i - 1) & 1].synchronize()
            done[(n - 1) & 1].synchronize()

        for ev_ in consumed:
            ev_.record(compute_stream)
        e2e_steps = max(3, min(args.steps, 10))
        run_e2e(3)
        barrier()
        t0 = time.perf_counter()
        run_e2e(e2e_steps)
        barrier()
        t_e2e = time.perf_counter() - t0
        e2e_note = ("the data loader's tensors (gt_class_ids, gt_boxes, full-size gt_masks, shuffle keys) copied from "
                    "pinned host memory every step (copy of step i+1 overlaps step i), target ROIs read back every step; "
                    "proposals, feature maps and head gradients are device-resident (the network produces them)")
        e2e_maps = "device-resident"

    # ---- max over ranks ----
    if world > 1:
        tt = torch.tensor([ms_total, t_e2e, ms_roof_avg, 0.0 if args.no_pipelined else ms_pipe, ms_eager], device=dev,
                          dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, t_e2e, ms_roof_avg, ms_eager = tt[0].item(), tt[1].item(), tt[2].item(), tt[4].item()
        if not args.no_pipelined:
            ms_pipe = tt[3].item()
    value = world * B * timed_steps / (ms_total * 1e-3)
    if not args.no_e2e:
        e2e = {"value": world * B * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "maps": e2e_maps, "note": e2e_note}
    if not args.no_pipelined:
        pipelined = {"value": world * B * timed_steps / (ms_pipe * 1e-3), "unit": UNIT, "streams": 2,
                     "ms_per_step": ms_pipe / timed_steps,
                     "note": "same K steps (eager layer calls) issued alternately on two CUDA streams, one workspace set "
                             "per stream"}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        if not training:
            # algorithmic bytes of the 7x7 launch: output written once + every sampled feature-map pixel read once +
            # the boxes (DESIGN.md "Measurement"); SURVEY 8(d)'s closed form is its upper bound (all pixels touched)
            rois_np = outs[0].cpu().numpy()
            _, roi_map7 = F.roialign_forward(outs[0], d["image_meta"], d_maps, (7, 7))
            out7 = B * P_ * 49 * 256 * 4
            bytes_roof = out7 + touched_map_bytes(rois_np, roi_map7.cpu().numpy(), S, 7, 7) + 16 * B * P_
            bytes_upper = B * algorithmic_bytes(S, P_, 7, 7)
            roof_kernel = "roialign_fwd_kernel<2, 1> (7x7, N=%d, +prep)" % P_
            roof_extra = {"survey_closed_form_bytes_per_launch": bytes_upper,
                          "real_rois_per_image": float((np.abs(rois_np).sum(-1) > 0).sum(1).mean())}
            stage_alg = stage_bytes(S, A, NC=NC)
        else:
            bytes_roof = B * backward_algorithmic_bytes(S, TRAIN_T, 14, 14)
            roof_kernel = "PyramidROIAlign backward 14x14, T=%d (deterministic: roialign_bwd_gather_kernel<2> + its count / alloc / fill passes)" % TRAIN_T
            roof_extra = {"definition": "SURVEY 8(d): gradient read once + every gradient-map pixel written once + boxes"}
            stage_alg = None
        achieved = bytes_roof / (ms_roof_avg * 1e-3) / 1e9
        default_shape = args.batch is None and args.img_size is None and args.num_classes is None
        step_traffic = measured_traffic(args.config, args.regime, "step_dram_bytes") if default_shape else None
        roof_traffic = measured_traffic(args.config, args.regime, "roofline_kernel_dram_bytes") if default_shape else None
        cpu = None
        if not args.no_cpu_baseline:
            if all_cpus:
                os.sched_setaffinity(0, all_cpus)     # the CPU baseline gets every host core
            ips, sec, reps, threads = time_cpu(x, cfg, c, args.cpu_seconds)
            cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{reps} x the full per-GPU batch of {B} images (whole stage), median; oracle/ C restatement "
                             "of the reference's TF CPU path with OpenMP"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": timed_steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / timed_steps, "higher_is_better": True,
            "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_object(args, c, world),
            "value_from": graph_note,
            "eager": {"value": world * B * timed_steps / (ms_eager * 1e-3), "ms_per_step": ms_eager / timed_steps,
                      "note": "the instrumented pass: eager layer calls through Python, CUDA events around the roofline "
                              "kernel's layer call in every step"},
            "anchors": A,
            "stage": {"closed_form_bytes_per_image": stage_alg,
                      "dram_bytes_per_step": step_traffic,
                      "dram_gbs": (world * step_traffic / (ms_total / timed_steps * 1e-3) / 1e9) if step_traffic else None,
                      "note": "dram_* = dram__bytes_read+write of every kernel of one step from the committed ncu capture "
                              "of this build (profiles/roofline_traffic.json); null when the capture is stale"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": KERNELS_PER_STEP * timed_steps,
            "roofline": dict({"bound": "hbm", "kernel": roof_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                              "frac": achieved / peak,
                              "traffic": roof_traffic,
                              "peak_source": peak_src, "ms_per_launch": ms_roof_avg,
                              "algorithmic_bytes_per_launch": bytes_roof,
                              "timed_in": "the instrumented (eager) pass of the same K steps"}, **roof_extra),
            "cpu_baseline": cpu,
            "pipelined": pipelined,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
