#!/usr/bin/env python
"""Vote kernel timing for a few launch geometries (reads DS_VOTE_* at every launch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deep_staple_b200 as ds  # noqa: E402

dev = torch.device("cuda:0")
T, N, S = 13, 30, (128, 128, 128)
V = S[0] * S[1] * S[2]
gen = torch.Generator(device=dev).manual_seed(4)
dps = 2 * torch.randn((T, N), generator=gen, device=dev)
which = sys.argv[1:] or ["u8", "i64"]
for name, dt, sz in (("u8", torch.uint8, 1), ("i64", torch.int64, 8)):
    if name not in which:
        continue
    labels = (torch.rand((T, N, *S), generator=gen, device=dev) < 0.3).to(dt)
    out = torch.empty((T, *S), dtype=dt, device=dev)
    alg = T * V * (N + 1) * sz
    for per_sm, waves in ((2, 8), (3, 8), (4, 4), (4, 16), (6, 8), (8, 8), (8, 16)):
        os.environ["DS_VOTE_CTAS_PER_SM"], os.environ["DS_VOTE_WAVES"] = str(per_sm), str(waves)
        sec = bench.timed(lambda: ds.consensus_batched(labels, dps, out=out), 20, 3)
        print(f"vote {name} per_sm={per_sm} waves={waves}: {sec / 20 * 1e6:8.1f} us  {alg * 20 / sec / 1e9:8.1f} GB/s", flush=True)
    del labels, out
