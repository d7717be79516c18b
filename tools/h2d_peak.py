"""Measured host->device copy bandwidth from page-locked memory (the ceiling of bench.py's e2e leg)."""
import torch
n = 128 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
for size in (1 << 20, 8 << 20, 14 << 20, 64 << 20, 128 << 20):
    with torch.cuda.stream(s):
        for _ in range(3):
            d[:size].copy_(h[:size], non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 20
        for _ in range(reps):
            d[:size].copy_(h[:size], non_blocking=True)
        e1.record()
    s.synchronize()
    print("%4d MiB chunks: %.1f GB/s" % (size >> 20, size * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9))
