#!/usr/bin/env python
"""uint8 vote: LDG kernel vs the staged (bulk-copy) kernel, cfg3's per-GPU shard and the full set."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deep_staple_b200 as ds  # noqa: E402

dev = torch.device("cuda:0")
N, S = 30, (128, 128, 128)
V = S[0] * S[1] * S[2]
gen = torch.Generator(device=dev).manual_seed(4)
for T in (13, 100):
    dps = 2 * torch.randn((T, N), generator=gen, device=dev)
    labels = (torch.rand((T, N, *S), generator=gen, device=dev) < 0.3).to(torch.uint8)
    alg = T * V * (N + 1)
    outs = {}
    for odt in (torch.uint8, torch.int64):
        out = torch.empty((T, *S), dtype=odt, device=dev)
        algo = T * V * (N + out.element_size())
        for staged, run in ((0, 1), (1, 1), (1, 2), (1, 4)):
            os.environ["DS_VOTE_STAGED"], os.environ["DS_VOTE_STG_RUN"] = str(staged), str(run)
            sec = bench.timed(lambda: ds.consensus_batched(labels, dps, out=out, out_dtype=odt), 20, 3)
            outs[(odt, staged, run)] = out.clone()
            same = torch.equal(outs[(odt, 0, 1)], out)
            print(f"T={T} out={str(odt)[6:]} staged={staged} run={run}: {sec / 20 * 1e6:8.1f} us  {algo * 20 / sec / 1e9:8.1f} GB/s  "
                  f"identical_to_ldg={same}", flush=True)
        del out
    os.environ["DS_VOTE_STAGED"] = "1"
    os.environ["DS_VOTE_STG_RUN"] = "1"
    _, ties1 = ds.consensus_batched(labels, dps, count_near_ties=True)
    os.environ["DS_VOTE_STAGED"] = "0"
    _, ties0 = ds.consensus_batched(labels, dps, count_near_ties=True)
    print(f"T={T} near ties staged={int(ties1)} ldg={int(ties0)}", flush=True)
    del labels, outs
