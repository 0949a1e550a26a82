"""Pinned host -> device copy rate of this box (the ceiling of the bench line's e2e figure): python scripts/h2d_ceiling.py"""
import torch

for mb in (64, 256, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"H2D {mb} MiB pinned: {10 * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
    e0.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"D2H {mb} MiB pinned: {10 * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
