"""Pinned host -> device copy bandwidth of the box (copy engine), for the e2e ceiling discussion in DESIGN.md section 5."""
import torch
dev = torch.device("cuda:0")
for mb in (16, 64, 256, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 8
    e0.record()
    for _ in range(n):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"H2D {mb:5d} MB: {n * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9:6.1f} GB/s", flush=True)
    e0.record()
    for _ in range(n):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"D2H {mb:5d} MB: {n * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9:6.1f} GB/s", flush=True)
