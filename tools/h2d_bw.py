"""Pinned host -> device copy bandwidth of this box (context for the e2e number)."""
import time, torch
dev = torch.device("cuda:0")
for mb in (16, 64, 256):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"H2D pinned {mb:4d} MiB: {ms:.3f} ms  {mb / 1024 / (ms / 1e3):.1f} GiB/s")
