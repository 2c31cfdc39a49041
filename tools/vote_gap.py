#!/usr/bin/env python
"""cfg1-sized consensus vote (1 target x N candidates x 128^3): public Python call vs raw C-ABI launches
(stream and CUDA graph) -- separates host overhead from the kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deep_staple_b200 as ds  # noqa: E402
from deep_staple_b200 import _binding as b  # noqa: E402

dev = torch.device("cuda:0")
lib = b.load()
T, N, V = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 10, 128 ** 3
for name, dt, code, sz in (("int64", torch.int64, b.DS_I64, 8), ("uint8", torch.uint8, b.DS_U8, 1)):
    nsets = max(2, min(16, int(1.5e9 // (T * N * V * sz))))
    g = torch.Generator(device=dev).manual_seed(1)
    sets = [((torch.rand((T, N, V), generator=g, device=dev) < 0.3).to(dt), torch.empty((T, V), dtype=dt, device=dev))
            for _ in range(nsets)]
    dps = (2 * torch.randn((T, N), generator=g, device=dev)).contiguous()
    st = {"i": 0}

    def api():
        lab, out = sets[st["i"] % nsets]
        st["i"] += 1
        ds.consensus_batched(lab.view(T, N, 128, 128, 128), dps, out=out.view(T, 128, 128, 128))

    def raw(stream):
        lab, out = sets[st["i"] % nsets]
        st["i"] += 1
        rc = lib.ds_consensus_vote(lab.data_ptr(), code, dps.data_ptr(), 1, T, N, V, out.data_ptr(), code, None, stream)
        assert rc == 0, lib.ds_last_error_string()

    cur = torch.cuda.current_stream(dev).cuda_stream
    t_api = bench.timed(api, 300, 10) / 300
    t_raw = bench.timed(lambda: raw(cur), 300, 10) / 300
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        raw(side.cuda_stream)
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(60):
                raw(side.cuda_stream)
    graph.replay()
    t_graph = bench.timed(graph.replay, 10, 2) / 600
    alg = T * V * (N + 1) * sz
    peak, _ = bench.peak_hbm()
    print(f"vote 1x{N}x128^3 {name}: python API {t_api * 1e6:6.1f} us | C ABI stream {t_raw * 1e6:6.1f} us | C ABI graph "
          f"{t_graph * 1e6:6.1f} us = {alg / t_graph / 1e9:6.0f} GB/s, {alg / t_graph / 1e9 / peak:.2f} of measured HBM")
