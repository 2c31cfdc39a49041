#!/usr/bin/env python
"""Kernel-only timing sweeps over the launch-geometry knobs (DS_LOSS_*/DS_VOTE_* environment
variables are read at every launch).  Prints one line per configuration.

    python tools/tune.py loss|vote [steps]
"""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def time_fn(fn, steps, warmup=5):
    return bench.timed(fn, steps, warmup)


def loss(steps):
    dev = torch.device("cuda:0")
    step = bench.LossStep(dev, 3, 0, 1)
    alg = step.B * step.V * 24
    for per_sm, waves in itertools.product((2, 3, 4, 5, 6, 8), (1, 2, 3, 4, 8)):
        os.environ["DS_LOSS_CTAS_PER_SM"], os.environ["DS_LOSS_WAVES"] = str(per_sm), str(waves)
        sec = time_fn(step.kernel, steps)
        print(f"loss per_sm={per_sm} waves={waves}: {sec / steps * 1e6:8.2f} us  {alg * steps / sec / 1e9:8.1f} GB/s", flush=True)


def variants(steps):
    """fused fwd+bwd vs gradient-only (pure stream, no statistics) vs statistics-only (out-of-line mode)"""
    import deep_staple_b200 as ds
    from deep_staple_b200 import _binding as b

    dev = torch.device("cuda:0")
    step = bench.LossStep(dev, 3, 0, 1)
    lib = b.load()
    V, B, C = step.V, step.B, step.C
    w = torch.rand(B, device=dev) + 0.5

    def bwd_only():
        k = step.i % 3
        step.i += 1
        s = step.sets[k]
        lib.ds_dp_loss_backward_logits(s["logits"].data_ptr(), b.DS_F32, s["target"].data_ptr(), b.DS_I64, B, C, V,
                                       w.data_ptr(), 1, step.scale.data_ptr(), step.dlogits[k].data_ptr(), step.stream)

    def fwd_only():
        k = step.i % 3
        step.i += 1
        s = step.sets[k]
        lib.ds_dp_loss_fwd_bwd(s["logits"].data_ptr(), b.DS_F32, s["target"].data_ptr(), b.DS_I64, None, B, C, V,
                               step.bare.data_ptr(), step.fixed.data_ptr(), B, 0, step.cw.data_ptr(), 1,
                               step.scale.data_ptr(), None, step.stats.data_ptr(), step.loss_out.data_ptr(),
                               step.d_bare.data_ptr(), step.per_inst.data_ptr(), step.ws.data_ptr(), step.ws.numel() * 8,
                               step.stream)

    geoms = ((4, 1), (3, 1), (3, 2), (4, 2), (6, 2), (3, 4))
    if os.environ.get("DS_TUNE_GEOM"):
        geoms = tuple(tuple(int(x) for x in g.split("x")) for g in os.environ["DS_TUNE_GEOM"].split(","))
    for per_sm, waves in geoms:
        os.environ["DS_LOSS_CTAS_PER_SM"], os.environ["DS_LOSS_WAVES"] = str(per_sm), str(waves)
        for name, fn, bytes_ in (("fwd+bwd", step.kernel, 24), ("bwd-only", bwd_only, 24), ("fwd-only", fwd_only, 16)):
            sec = time_fn(fn, steps)
            print(f"{name:9s} per_sm={per_sm} waves={waves}: {sec / steps * 1e6:8.2f} us  {B * V * bytes_ * steps / sec / 1e9:8.1f} GB/s", flush=True)


def vote(steps):
    import deep_staple_b200 as ds

    dev = torch.device("cuda:0")
    T, N, S = 13, 30, (128, 128, 128)
    V = S[0] * S[1] * S[2]
    gen = torch.Generator(device=dev).manual_seed(4)
    dps = 2 * torch.randn((T, N), generator=gen, device=dev)
    for name, dt, sz in (("u8", torch.uint8, 1), ("i64", torch.int64, 8)):
        labels = (torch.rand((T, N, *S), generator=gen, device=dev) < 0.3).to(dt)
        out = torch.empty((T, *S), dtype=dt, device=dev)
        alg = T * V * (N + 1) * sz
        for per_sm, waves in itertools.product((2, 3, 4, 6, 8), (1, 2, 4, 8, 16)):
            os.environ["DS_VOTE_CTAS_PER_SM"], os.environ["DS_VOTE_WAVES"] = str(per_sm), str(waves)
            sec = time_fn(lambda: ds.consensus_batched(labels, dps, out=out), steps, 3)
            print(f"vote {name} per_sm={per_sm} waves={waves}: {sec / steps * 1e6:8.1f} us  {alg * steps / sec / 1e9:8.1f} GB/s", flush=True)
        del labels, out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "loss"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    {"loss": loss, "vote": vote, "variants": variants}[which](steps)
