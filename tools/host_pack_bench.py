#!/usr/bin/env python
"""Host label packer (ds_host_narrow_labels): GB/s of int64 read per thread count, pageable and pinned buffers."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_staple_b200 as ds  # noqa: E402

print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
x = torch.randint(0, 2, (8, 128, 128, 128), dtype=torch.int64)
for pinned in (False, True):
    src = x.pin_memory() if pinned else x
    out = torch.empty(x.shape, dtype=torch.uint8)
    out = out.pin_memory() if pinned else out
    for th in (1, 2, 4, 8, 16, 32):
        ds.narrow_labels_host(src, out, th)
        t = time.perf_counter()
        for _ in range(10):
            ds.narrow_labels_host(src, out, th)
        dt = (time.perf_counter() - t) / 10
        print(f"pinned={pinned} threads={th:2d}: {dt * 1e3:6.2f} ms  {x.numel() * 8 / dt / 1e9:6.1f} GB/s read", flush=True)
assert torch.equal(out, x.to(torch.uint8))
