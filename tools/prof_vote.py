#!/usr/bin/env python
"""A few launches of the consensus vote on a cfg3 shard (for ncu).  usage: prof_vote.py <i64|u8> [T] [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_staple_b200 as ds  # noqa: E402

dt = {"i64": torch.int64, "u8": torch.uint8}[sys.argv[1] if len(sys.argv) > 1 else "i64"]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 13
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
lab = (torch.rand((T, N, 128, 128, 128), generator=g, device=dev) < 0.3).to(dt)
dps = 2 * torch.randn((T, N), generator=g, device=dev)
out = torch.empty((T, 128, 128, 128), dtype=dt, device=dev)
for _ in range(4):
    ds.consensus_batched(lab, dps, out=out)
torch.cuda.synchronize()
print("done", int(out.sum()))
