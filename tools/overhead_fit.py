#!/usr/bin/env python
"""Launch-time floor of the fused loss kernel: time vs volume at fixed batch -> fixed cost + per-byte cost."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deep_staple_b200 import _binding as b  # noqa: E402

dev = torch.device("cuda:0")
lib = b.load()
B, C = 8, 2
for tma in (0, 1):
    os.environ["DS_LOSS_TMA"] = str(tma)
    pts = []
    for logv in (12, 14, 16, 18, 19, 20, 21):
        V = 1 << logv
        nsets = max(2, min(64, int(2e9 // (B * V * 24))))
        sets = [(torch.randn((B, C, V), device=dev), (torch.rand((B, V), device=dev) < 0.3).long(),
                 torch.empty((B, C, V), device=dev)) for _ in range(nsets)]
        bare = torch.randn(B, device=dev)
        cw = torch.ones(C, device=dev)
        scale = torch.tensor(1024.0, device=dev)
        stats = torch.zeros((B, C, 4), dtype=torch.float64, device=dev)
        loss_out = torch.zeros(2, device=dev)
        d_bare = torch.zeros(B, device=dev)
        per_inst = torch.zeros((B, 4), device=dev)
        ws = torch.zeros((int(lib.ds_dp_loss_workspace_bytes(B, C)) + 7) // 8, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        st = {"i": 0}

        def fn():
            z, t, g = sets[st["i"] % nsets]
            st["i"] += 1
            lib.ds_dp_loss_fwd_bwd(z.data_ptr(), b.DS_F32, t.data_ptr(), b.DS_I64, None, B, C, V, bare.data_ptr(), None, B, 0,
                                   cw.data_ptr(), 1, scale.data_ptr(), g.data_ptr(), stats.data_ptr(), loss_out.data_ptr(),
                                   d_bare.data_ptr(), per_inst.data_ptr(), ws.data_ptr(), ws.numel() * 8, stream)

        sec = bench.timed(fn, 300, 10) / 300
        pts.append((B * V * 24, sec))
        print(f"tma={tma} V=2^{logv:2d} bytes={B * V * 24 / 1e6:8.2f} MB  {sec * 1e6:8.2f} us  {B * V * 24 / sec / 1e9:8.1f} GB/s", flush=True)
        del sets
    (x0, y0), (x1, y1) = pts[-3], pts[-1]
    slope = (y1 - y0) / (x1 - x0)
    print(f"tma={tma}: marginal {1 / slope / 1e9:.0f} GB/s, intercept {(y1 - slope * x1) * 1e6:.1f} us")
