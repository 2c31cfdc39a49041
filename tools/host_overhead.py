#!/usr/bin/env python
"""Host-side cost of the public Python API per call (the kernel of a tiny problem is faster than the host, so the
wall time per call of an unsynchronised loop IS the host overhead), with a cProfile breakdown.

    python tools/host_overhead.py [--profile]
"""
import cProfile
import os
import pstats
import sys
import time

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_staple_b200 as ds  # noqa: E402

dev = torch.device("cuda:0")
B, C, S, M = 8, 2, (16, 16, 16), 600
gen = torch.Generator(device=dev).manual_seed(0)
logits = 2 * torch.randn((B, C, *S), generator=gen, device=dev)
target = (torch.rand((B, *S), generator=gen, device=dev) < 0.05).long()
emb = nn.Embedding(M, 1, sparse=True).to(dev)
opt = torch.optim.SparseAdam(emb.parameters(), lr=0.1)
idx = torch.randperm(M, device=dev)[:B]
fixed = torch.rand(M, device=dev) + 1
cw = torch.tensor([0.6, 1.4], device=dev)
scaler = torch.amp.GradScaler("cuda", init_scale=65536.0)
scaler.scale(torch.zeros((), device=dev))  # creates the scale tensor


def inline_step():  # in-line mode: gradient into the logits and the data parameters, autograd backward
    z = logits.detach().requires_grad_(True)
    loss = ds.dp_weighted_loss(z, target, emb(idx).squeeze(), fixed[idx], True, grad_scaler=scaler)
    scaler.scale(loss).backward()
    emb.weight.grad = None


def ool_step():  # out-of-line mode (the reference's default): no gradient into the logits
    loss = ds.dp_weighted_loss(logits, target, emb(idx).squeeze(), fixed[idx], True, grad_scaler=scaler)
    scaler.scale(loss).backward()
    emb.weight.grad = None


def fwd_only():
    ds.dp_weighted_loss(logits, target, emb.weight.detach()[idx].squeeze(), fixed[idx], True, grad_scaler=scaler)


def ce_step():
    z = logits.detach().requires_grad_(True)
    loss = ds.weighted_ce(z, target, cw, grad_scaler=scaler, class_counts=counts)
    scaler.scale(loss).backward()


counts = torch.bincount(target.reshape(-1), minlength=C).double()
cases = {"dp_weighted_loss in-line + backward": inline_step, "dp_weighted_loss out-of-line + backward": ool_step,
         "dp_weighted_loss forward only (no autograd)": fwd_only, "weighted_ce + backward": ce_step}
if hasattr(ds, "LossSection"):
    section = ds.LossSection(emb, opt, scaler, fixed_weighting=fixed, use_risk_regularization=True)
    cases["LossSection out-of-line: loss + optimiser step + scaler update"] = lambda: section(logits, target, idx)

for name, fn in cases.items():
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name:70s} {1e6 * (t1 - t0) / n:7.1f} us host per call", flush=True)

if "--profile" in sys.argv:
    for name in ("weighted_ce + backward", "dp_weighted_loss out-of-line + backward"):
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(2000):
            cases[name]()
        pr.disable()
        torch.cuda.synchronize()
        print("==", name)
        pstats.Stats(pr).sort_stats("tottime").print_stats(25)
