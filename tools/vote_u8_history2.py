#!/usr/bin/env python
"""Which buffer's placement slows the staged uint8 vote down after a 50 GB allocate / free cycle?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deep_staple_b200 as ds  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
T, N, S = 100, 30, (128, 128, 128)
g = torch.Generator(device=dev).manual_seed(4)
dps = 2 * torch.randn((T, N), generator=g, device=dev)


def t(labels, res, tag):
    sec = bench.timed(lambda: ds.consensus_batched(labels, dps, out=res), 20, 3)
    print(f"{tag:64s} {sec / 20 * 1e3:.3f} ms  labels {labels.data_ptr():#x} out {res.data_ptr():#x}", flush=True)


lab = torch.cat([bench.ellipsoid_labels(1, N, S, dev, 1000 + k, torch.uint8) for k in range(T)])
res_old = torch.empty((T, *S), dtype=torch.uint8, device=dev)
t(lab, res_old, "before: old labels, old out")
big = torch.empty(50 * 1024 ** 3, dtype=torch.uint8, device=dev)
big.zero_()
torch.cuda.synchronize()
del big
torch.cuda.empty_cache()
t(lab, res_old, "after 50 GB alloc/free: old labels, old out")
res_new = torch.empty((T, *S), dtype=torch.uint8, device=dev)
t(lab, res_new, "old labels, NEW out")
lab_new = lab.clone()
t(lab_new, res_new, "NEW labels, NEW out")
t(lab_new, res_old, "NEW labels, old out")
os.environ["DS_VOTE_STAGED"] = "0"
t(lab, res_new, "LDG kernel: old labels, NEW out")
t(lab, res_old, "LDG kernel: old labels, old out")
os.environ["DS_VOTE_STAGED"] = "1"
for off in (256, 4096, 65536, 1 << 20):
    pad = torch.empty(off, dtype=torch.uint8, device=dev)
    r = torch.empty((T, *S), dtype=torch.uint8, device=dev)
    t(lab, r, f"old labels, another new out (after a {off} B pad)")
print(torch.cuda.memory_summary(abbreviated=True)[:600])
