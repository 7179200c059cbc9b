#!/usr/bin/env python
"""Raw pinned host<->device copy bandwidth of this box (the ceiling of bench.py's e2e number)."""
import torch

if __name__ == "__main__":
    n = 360_000_000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("%s %.1f GB/s" % (name, 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
