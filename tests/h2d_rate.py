"""Pinned host -> device copy rate on this box (the e2e bound of bench.py: 192 MiB per step)."""
import torch
x = torch.empty(192 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    d.copy_(x, non_blocking=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"H2D 192 MiB pinned: {ms:.3f} ms  {(192 << 20) / ms / 1e6:.1f} GB/s")
