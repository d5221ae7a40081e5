"""Experiment: device->host copy bandwidth into pinned memory (the ceiling of the host-blob path)."""
import torch

for mb in (64, 176, 403):
    n = mb * 1000 * 1000
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"D2H {mb} MB: {ms:.3f} ms  {n / ms * 1e-6:.1f} GB/s")
