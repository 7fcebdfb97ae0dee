"""Batch sharding across the GPUs of one box: one process per GPU, contiguous image ranges, no collective on the
hot path.  The only exchange is the optional gather of the fixed-size detection tensors
for host-side metrics, after the stage has finished."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world_size):
    """Contiguous [start, stop) of `total` images for `rank`; earlier ranks take the remainder."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(int(total), world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(total, world_size):
    return [shard_range(total, r, world_size)[1] - shard_range(total, r, world_size)[0] for r in range(world_size)]


def gather_detections(local, total, group=None):
    """All-gather per-rank detections [b_r, D, 6] into [total, D, 6] (every rank gets the full tensor).
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
