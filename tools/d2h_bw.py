#!/usr/bin/env python
"""Device-to-host copy bandwidth into pinned memory by transfer size (what bounds bench.py's e2e)."""
import time
import torch
for mb in (6.2208, 12.4416, 24.8832, 49.7664, 99.5328, 256):
    n = int(mb * 1e6)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(3)]
    s = torch.cuda.Stream()
    reps = max(8, int(2e9 / n))
    with torch.cuda.stream(s):
        for i in range(4):
            hs[i % 3].copy_(d, non_blocking=True)
        s.synchronize()
        t0 = time.perf_counter()
        for i in range(reps):
            hs[i % 3].copy_(d, non_blocking=True)
        s.synchronize()
        dt = time.perf_counter() - t0
    print("%8.2f MB per copy: %.1f GB/s" % (mb, n * reps / dt / 1e9), flush=True)
