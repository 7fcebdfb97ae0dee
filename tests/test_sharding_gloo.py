"""world_size-2 gloo run of the N>1 path on CPU: contiguous image shards, and the host-side gather of the
fixed-size detection tensors (the only exchange; nothing on the hot path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from maskrcnn_tf2_b200.sharding import gather_detections, shard_range
    from maskrcnn_tf2_b200 import synth
    start, stop = shard_range(total, rank, world)
    # each rank computes the detections of its own images (CPU oracle stands in for the device here)
    rng = np.random.default_rng(9)
    rois = rng.uniform(0, 1, (total, 50, 4)).astype(np.float32)
    rois[..., 2:] = np.minimum(rois[..., :2] + 0.2, 1.0)
    probs, deltas = synth.head_outputs(rng, total, 50, 5)
    meta = synth.image_meta(total, 256, 5)
    local = oracle.detection_layer(rois[start:stop], probs[start:stop], deltas[start:stop], meta[start:stop],
                                   [0.1, 0.1, 0.2, 0.2], 0.5, 10, 0.3)["detections"]
    full = gather_detections(torch.from_numpy(local), total)
    ref = oracle.detection_layer(rois, probs, deltas, meta, [0.1, 0.1, 0.2, 0.2], 0.5, 10, 0.3)["detections"]
    ok = tuple(full.shape) == (total, 10, 6) and np.array_equal(full.numpy(), ref)
    with open(os.path.join(tmpdir, f"rank{rank}.txt"), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_and_detection_gather(tmp_path):
    world, total = 2, 5          # uneven split: 3 + 2 images
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"
