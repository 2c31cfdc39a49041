#!/usr/bin/env python
"""End-to-end consensus from a train_label_snapshot (consensus.ipynb:353-418 minus STAPLE): the CPU oracle
(reference lines on torch CPU) vs deep_staple_b200 (sparse-COO ingestion, vote and Dice on the GPU)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deep_staple_b200 as ds  # noqa: E402
from oracle import consensus as oc  # noqa: E402

T, N, S = int(os.environ.get("T", 6)), int(os.environ.get("N", 10)), (128, 128, 128)
dev = torch.device("cuda:0")
labels = bench.ellipsoid_labels(T, N, S, dev, 0, torch.int64).cpu()
expert = bench.ellipsoid_labels(T, 1, S, dev, 99, torch.int64).cpu()
ids, mods, labs, preds = [], [], [], []
for t in range(T):
    for n in range(N):
        ids.append(f"{100 + t:03d}l:m{200 + n:03d}r")
        mods.append(labels[t, n])
        labs.append(expert[t, 0])
        preds.append(expert[t, 0].unsqueeze(0))
snap = ds.build_snapshot(torch.randn(T * N), ids, list(range(T * N)), labels=labs, modified_labels=mods, train_predictions=preds)
nnz = snap["modified_labels"]._nnz()
print(f"snapshot: {T} targets x {N} atlases at 128^3, {nnz} stored voxels ({100.0 * nnz / (T * N * S[0] ** 3):.2f} % foreground)")

torch.set_num_threads(os.cpu_count())
t0 = time.perf_counter()
ref = oc.consensus_from_snapshot(dict(snap))
t_ref = time.perf_counter() - t0

ds.consensus_from_snapshot(dict(snap))  # warm-up (library load, allocator)
torch.cuda.synchronize()
t0 = time.perf_counter()
ours = ds.consensus_from_snapshot(dict(snap))
torch.cuda.synchronize()
t_ours = time.perf_counter() - t0
for k in ref:
    assert torch.equal(ours[k]["dp_consensus"].to_dense(), ref[k]["dp_consensus"].to_dense()), k
print(f"oracle (torch CPU, {os.cpu_count()} threads): {t_ref * 1e3:9.1f} ms   ({T * N * S[0] ** 3 / t_ref / 1e9:.3f} Gvoxel-cand/s)")
print(f"deep_staple_b200 (host loop + sparse vote kernel + Dice kernel, CPU tensors in/out): {t_ours * 1e3:9.1f} ms   "
      f"({T * N * S[0] ** 3 / t_ours / 1e9:.3f} Gvoxel-cand/s)   speed-up {t_ref / t_ours:.1f}x, maps identical")

# kernel-level: sparse vote vs dense vote for one target
lb = [snap["modified_labels"][i] for i in range(N)]
dps = [snap["data_parameters"][i] for i in range(N)]
dense = torch.stack([l.to_dense() for l in lb]).to(dev).unsqueeze(0)
w = torch.stack([d.reshape(()) for d in dps]).to(dev).reshape(1, -1)
sec_dense = bench.timed(lambda: ds.consensus_batched(dense, w), 50, 3) / 50
lb_dev = [l.to(dev) for l in lb]
sec_sparse = bench.timed(lambda: ds.consensus_sparse(lb_dev, dps), 20, 3) / 20
print(f"one target, {N} candidates: dense int64 kernel {sec_dense * 1e6:.1f} us; sparse path incl. index flattening {sec_sparse * 1e6:.1f} us")
