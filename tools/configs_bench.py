#!/usr/bin/env python
"""Kernel-level timing of every BASELINE.json configuration (and dtype variants) through the C ABI.
Prints one line per case: launch time, algorithmic GB/s, fraction of the measured HBM peak."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deep_staple_b200 as ds  # noqa: E402
from deep_staple_b200 import _binding as b  # noqa: E402

dev = torch.device("cuda:0")
lib = b.load()
peak, _ = bench.peak_hbm()
DT = {torch.float32: b.DS_F32, torch.float16: b.DS_F16, torch.bfloat16: b.DS_BF16, torch.int64: b.DS_I64, torch.uint8: b.DS_U8}


def loss_case(name, B, C, S, ldt, tdt, grad=True, nsets=3, steps=200):
    V = math.prod(S)
    sets = []
    for i in range(nsets):
        g = torch.Generator(device=dev).manual_seed(i)
        z = (2 * torch.randn((B, C, *S), generator=g, device=dev)).to(ldt)
        t = (torch.rand((B, *S), generator=g, device=dev) * C * 0.35).long().clamp_(0, C - 1).to(tdt)
        sets.append((z, t, torch.empty_like(z) if grad else None))
    bare = torch.randn(B, device=dev)
    fixed = torch.rand(B, device=dev) + 1
    cw = torch.ones(C, device=dev)
    scale = torch.tensor(1024.0, device=dev)
    stats = torch.zeros((B, C, 4), dtype=torch.float64, device=dev)
    loss_out = torch.zeros(2, device=dev)
    d_bare = torch.zeros(B, device=dev)
    per_inst = torch.zeros((B, 4), device=dev)
    ws = torch.zeros((int(lib.ds_dp_loss_workspace_bytes(B, C)) + 7) // 8, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    state = {"i": 0}

    def fn():
        z, t, g = sets[state["i"] % nsets]
        state["i"] += 1
        rc = lib.ds_dp_loss_fwd_bwd(z.data_ptr(), DT[ldt], t.data_ptr(), DT[tdt], None, B, C, V, bare.data_ptr(),
                                    fixed.data_ptr(), B, 0, cw.data_ptr(), 1, scale.data_ptr(),
                                    g.data_ptr() if grad else None, stats.data_ptr(), loss_out.data_ptr(),
                                    d_bare.data_ptr(), per_inst.data_ptr(), ws.data_ptr(), ws.numel() * 8, stream)
        assert rc == 0, lib.ds_last_error_string()

    sec = bench.timed(fn, steps, 5) / steps
    sl, st = z.element_size(), t.element_size()
    alg = B * V * ((2 if grad else 1) * C * sl + st)
    ws_mb = nsets * B * V * ((2 if grad else 1) * C * sl + st) / 1e6
    print(f"{name:44s} {sec * 1e6:9.1f} us  {alg / sec / 1e9:8.1f} GB/s  frac {alg / sec / 1e9 / peak:5.3f}  "
          f"({B * V / sec / 1e9:7.1f} Gvox/s, rotating {ws_mb:.0f} MB)", flush=True)


def ce_case(name, B, C, S, ldt, counts_known, nsets=3, steps=100):
    """Row a1 (main_deep_staple.py:716-719): class-weighted CE value + dlogits through the public API
    (ds.weighted_ce + backward).  Without `class_counts` the op first histograms the labels (one extra read of
    the labels: 8 B per voxel on top of the 24), with them (e.g. from the data loader) it is one launch."""
    V = math.prod(S)
    sets = []
    for i in range(nsets):
        g = torch.Generator(device=dev).manual_seed(i)
        z = (2 * torch.randn((B, C, *S), generator=g, device=dev)).to(ldt).requires_grad_(True)
        t = (torch.rand((B, *S), generator=g, device=dev) < 0.05).long()
        sets.append((z, t, torch.bincount(t.reshape(-1), minlength=C).double()))
    cw = torch.tensor([0.6, 1.4] + [1.0] * (C - 2), device=dev)[:C]
    state = {"i": 0}

    def fn():
        z, t, counts = sets[state["i"] % nsets]
        state["i"] += 1
        loss = ds.weighted_ce(z, t, cw, grad_scaler=1024.0, class_counts=counts if counts_known else None)
        (loss * 1024.0).backward()
        z.grad = None

    sec = bench.timed(fn, steps, 5) / steps
    alg = B * V * (2 * C * sets[0][0].element_size() + 8 + (0 if counts_known else 8))
    # the kernel(s) alone through the C ABI (what the GPU does; the Python layer above is host-bound at this size)
    cwf = cw.float().contiguous()
    stats = torch.zeros((B, C, 4), dtype=torch.float64, device=dev)
    loss_out = torch.zeros(2, device=dev)
    scale = torch.tensor(1024.0, device=dev)
    ws = torch.zeros((int(lib.ds_dp_loss_workspace_bytes(B, C)) + 7) // 8, dtype=torch.int64, device=dev)
    grads = [torch.empty_like(z.detach()) for z, _, _ in sets]
    stream = torch.cuda.current_stream(dev).cuda_stream
    lstats = torch.zeros((B, C, 4), dtype=torch.float64, device=dev)

    def raw():
        k = state["i"] % nsets
        z, t, counts = sets[k]
        state["i"] += 1
        if not counts_known:
            rc = lib.ds_label_stats(t.data_ptr(), b.DS_I64, None, b.DS_I64, B, C, V, lstats.data_ptr(), ws.data_ptr(),
                                    ws.numel() * 8, stream)
            assert rc == 0, lib.ds_last_error_string()
        rc = lib.ds_weighted_ce_fwd_bwd(z.data_ptr(), DT[ldt], t.data_ptr(), b.DS_I64, B, C, V, cwf.data_ptr(),
                                        counts.data_ptr(), scale.data_ptr(), grads[k].data_ptr(), stats.data_ptr(),
                                        loss_out.data_ptr(), ws.data_ptr(), ws.numel() * 8, stream)
        assert rc == 0, lib.ds_last_error_string()

    ksec = bench.timed(raw, steps, 5) / steps
    print(f"{name:44s} {ksec * 1e6:9.1f} us  {alg / ksec / 1e9:8.1f} GB/s  frac {alg / ksec / 1e9 / peak:5.3f}  "
          f"({B * V / ksec / 1e9:7.1f} Gvox/s)  [C ABI; public API incl. autograd, host-bound: {sec * 1e6:.0f} us]", flush=True)


def vote_case(name, T, N, S, dt, steps=20, nsets=1):
    V = math.prod(S)
    g = torch.Generator(device=dev).manual_seed(1)
    sets = [((torch.rand((T, N, *S), generator=g, device=dev) < 0.3).to(dt), torch.empty((T, *S), dtype=dt, device=dev))
            for _ in range(nsets)]
    dps = 2 * torch.randn((T, N), generator=g, device=dev)
    state = {"i": 0}

    def fn():
        lab, out = sets[state["i"] % nsets]
        state["i"] += 1
        ds.consensus_batched(lab, dps, out=out)

    sec = bench.timed(fn, steps, 3) / steps
    sz = sets[0][0].element_size()
    alg = T * V * (N + 1) * sz
    print(f"{name:44s} {sec * 1e6:9.1f} us  {alg / sec / 1e9:8.1f} GB/s  frac {alg / sec / 1e9 / peak:5.3f}  "
          f"({T * N * V / sec / 1e9:7.1f} Gvox-cand/s)", flush=True)


def aten_cuda_case(name, B, C, S, ldt, steps=10):
    """Secondary baseline: the reference's own lines (main_deep_staple.py:738-762) on CUDA tensors, i.e. the
    stock ATen kernel composition on the same B200 (fwd + backward into logits and data parameters)."""
    from torch import nn

    V = math.prod(S)
    g = torch.Generator(device=dev).manual_seed(0)
    z0 = (2 * torch.randn((B, C, *S), generator=g, device=dev)).to(ldt)
    t = (torch.rand((B, *S), generator=g, device=dev) * C * 0.35).long().clamp_(0, C - 1)
    emb = nn.Embedding(600, 1, sparse=True).to(dev)
    idx = torch.arange(B, device=dev)
    fixed = torch.rand(600, device=dev) + 1
    n_dims = tuple(range(-len(S), 0))

    def fn():
        z = z0.detach().requires_grad_(True)
        zz = z.float()
        dp_loss = nn.CrossEntropyLoss(reduction="none")(zz, t).mean(n_dims)
        weight = torch.sigmoid(emb(idx).squeeze())
        weight = weight / weight.mean()
        weight = weight / fixed[idx]
        p_pred_num = (zz.argmax(1) > 0).sum(dim=n_dims).detach()
        risk = -weight * p_pred_num / V
        loss = (dp_loss * weight).sum() + risk.sum()
        (loss * 1024.0).backward()
        emb.weight.grad = None

    sec = bench.timed(fn, steps, 3) / steps
    alg = B * V * (2 * C * z0.element_size() + 8)
    print(f"{name:44s} {sec * 1e6:9.1f} us  {alg / sec / 1e9:8.1f} GB/s  frac {alg / sec / 1e9 / peak:5.3f}  "
          f"({B * V / sec / 1e9:7.1f} Gvox/s)  [stock torch CUDA composition]", flush=True)


_FILTER = [f for f in os.environ.get("DS_CASES", "").split(",") if f]


def _wrap(fn):
    def run(name, *a, **k):
        if _FILTER and not any(f in name for f in _FILTER):
            return
        fn(name, *a, **k)
    return run


if __name__ == "__main__":
    # DS_CASES=substr[,substr...] runs only the matching cases
    loss_case, vote_case, aten_cuda_case, ce_case = _wrap(loss_case), _wrap(vote_case), _wrap(aten_cuda_case), _wrap(ce_case)
    S3 = (128, 128, 128)
    loss_case("cfg2 fp32/int64 fwd+bwd", 8, 2, S3, torch.float32, torch.int64)
    loss_case("cfg2 fp16/int64 fwd+bwd (AMP)", 8, 2, S3, torch.float16, torch.int64)
    loss_case("cfg2 bf16/int64 fwd+bwd", 8, 2, S3, torch.bfloat16, torch.int64)
    loss_case("cfg2 fp32/uint8 fwd+bwd", 8, 2, S3, torch.float32, torch.uint8)
    loss_case("cfg2 fp16/uint8 fwd+bwd", 8, 2, S3, torch.float16, torch.uint8)
    loss_case("cfg2 fp32/int64 fwd only (out-of-line)", 8, 2, S3, torch.float32, torch.int64, grad=False)
    loss_case("cfg2 fp16/int64 fwd only (out-of-line)", 8, 2, S3, torch.float16, torch.int64, grad=False)
    loss_case("cfg2 fp32/uint8 fwd only", 8, 2, S3, torch.float32, torch.uint8, grad=False)
    loss_case("cfg2 fp16/uint8 fwd only", 8, 2, S3, torch.float16, torch.uint8, grad=False)
    loss_case("cfg4 2D 64x2x256x256 fp32/int64", 64, 2, (256, 256), torch.float32, torch.int64, nsets=8, steps=400)
    loss_case("cfg4 2D 64x2x256x256 fp16/int64", 64, 2, (256, 256), torch.float16, torch.int64, nsets=12, steps=400)
    loss_case("cfg5 8x3x192x160x160 fp32/int64", 8, 3, (192, 160, 160), torch.float32, torch.int64, nsets=2, steps=50)
    loss_case("cfg5 8x3x192x160x160 fp32/int64 fwd only", 8, 3, (192, 160, 160), torch.float32, torch.int64, grad=False, nsets=2, steps=50)
    loss_case("cfg5 8x3x192x160x160 fp16/int64", 8, 3, (192, 160, 160), torch.float16, torch.int64, nsets=2, steps=50)
    loss_case("cfg5 1x3x192x160x160 fp32/int64 (B=1)", 1, 3, (192, 160, 160), torch.float32, torch.int64, nsets=3)
    loss_case("4-class 8x4x128^3 fp32/int64", 8, 4, S3, torch.float32, torch.int64, nsets=2, steps=100)
    ce_case("a1 weighted CE cfg2 fp32, counts given", 8, 2, S3, torch.float32, True)
    ce_case("a1 weighted CE cfg2 fp32, label pre-pass", 8, 2, S3, torch.float32, False)
    ce_case("a1 weighted CE cfg2 fp16, counts given", 8, 2, S3, torch.float16, True)
    aten_cuda_case("ATen-on-B200 cfg2 fp32/int64 fwd+bwd", 8, 2, S3, torch.float32)
    aten_cuda_case("ATen-on-B200 cfg2 fp16/int64 fwd+bwd", 8, 2, S3, torch.float16)
    vote_case("cfg1 vote 1x10x128^3 int64", 1, 10, S3, torch.int64, steps=200, nsets=4)
    vote_case("cfg1 vote 1x10x128^3 uint8", 1, 10, S3, torch.uint8, steps=200, nsets=12)
    vote_case("cfg3 shard 13x30x128^3 int64", 13, 30, S3, torch.int64)
    vote_case("cfg3 shard 13x30x128^3 uint8", 13, 30, S3, torch.uint8)
