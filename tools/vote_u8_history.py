#!/usr/bin/env python
"""Does the staged uint8 vote depend on what ran before it in the process?  (bench.py's default run showed 1.17 ms
for the full set where the stand-alone run shows 0.97 ms.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deep_staple_b200 as ds  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
T, N, S = 100, 30, (128, 128, 128)


def make(kind):
    if kind == "ellipsoid":
        return torch.cat([bench.ellipsoid_labels(1, N, S, dev, 1000 + t, torch.uint8) for t in range(T)])
    g = torch.Generator(device=dev).manual_seed(4)
    return (torch.rand((T, N, *S), generator=g, device=dev) < 0.3).to(torch.uint8)


def vote_ms(labels, tag):
    g = torch.Generator(device=dev).manual_seed(4)
    dps = 2 * torch.randn((T, N), generator=g, device=dev)
    res = torch.empty((T, *S), dtype=torch.uint8, device=dev)
    sec = bench.timed(lambda: ds.consensus_batched(labels, dps, out=res), 20, 3)
    print(f"{tag:60s} {sec / 20 * 1e3:.3f} ms  (labels at {labels.data_ptr():#x}, fg {float(labels[:2].float().mean()):.4f})", flush=True)


for kind in ("random", "ellipsoid"):
    lab = make(kind)
    vote_ms(lab, f"{kind} labels, fresh process")
    del lab
    torch.cuda.empty_cache()
step = bench.LossStep(dev, 3, 0, 1)
sec = bench.timed(step.kernel, 200, 5)
print(f"loss steps: {sec / 200 * 1e6:.1f} us", flush=True)
lab = make("ellipsoid")
vote_ms(lab, "ellipsoid labels, after loss steps (LossStep alive)")
del step
torch.cuda.empty_cache()
vote_ms(lab, "ellipsoid labels, LossStep freed")
lab64 = lab.long()
res64 = torch.empty((T, *S), dtype=torch.int64, device=dev)
g = torch.Generator(device=dev).manual_seed(4)
dps = 2 * torch.randn((T, N), generator=g, device=dev)
sec = bench.timed(lambda: ds.consensus_batched(lab64, dps, out=res64), 5, 2)
print(f"int64 vote: {sec / 5 * 1e3:.3f} ms", flush=True)
del lab64, res64
torch.cuda.empty_cache()
vote_ms(lab, "ellipsoid labels, after the int64 vote")
