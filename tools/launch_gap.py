#!/usr/bin/env python
"""How much of a step is the gap BETWEEN launches?  cfg2 (or any B C D H W) fused loss launches timed
(a) back to back from the host on one stream, (b) as one CUDA graph of the same launches."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_staple_b200 import _binding as b  # noqa: E402

args = [int(x) for x in sys.argv[1:6]] if len(sys.argv) > 5 else [8, 2, 128, 128, 128]
B, C, S = args[0], args[1], tuple(args[2:5])
V = math.prod(S)
dev = torch.device("cuda:0")
lib = b.load()
nsets = max(3, int(1.3e9 // (B * V * 24)))
sets = []
for i in range(nsets):
    g = torch.Generator(device=dev).manual_seed(i)
    sets.append((2 * torch.randn((B, C, V), generator=g, device=dev),
                 (torch.rand((B, V), generator=g, device=dev) < 0.05).long(), torch.empty((B, C, V), device=dev)))
bare = torch.randn(B, device=dev); fixed = torch.rand(B, device=dev) + 1; cw = torch.ones(C, device=dev)
scale = torch.tensor(1024.0, device=dev); stats = torch.zeros((B, C, 4), dtype=torch.float64, device=dev)
loss_out = torch.zeros(2, device=dev); d_bare = torch.zeros(B, device=dev); per_inst = torch.zeros((B, 4), device=dev)
ws = torch.zeros((int(lib.ds_dp_loss_workspace_bytes(B, C)) + 7) // 8, dtype=torch.int64, device=dev)


def launch(i, stream):
    z, t, g = sets[i % nsets]
    rc = lib.ds_dp_loss_fwd_bwd(z.data_ptr(), b.DS_F32, t.data_ptr(), b.DS_I64, None, B, C, V, bare.data_ptr(),
                                fixed.data_ptr(), B, 0, cw.data_ptr(), 1, scale.data_ptr(), g.data_ptr(),
                                stats.data_ptr(), loss_out.data_ptr(), d_bare.data_ptr(), per_inst.data_ptr(),
                                ws.data_ptr(), ws.numel() * 8, stream)
    assert rc == 0, lib.ds_last_error_string()


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


K = 60
st = torch.cuda.current_stream(dev).cuda_stream
for i in range(20):
    launch(i, st)
us_stream = timed(lambda: [launch(i, st) for i in range(K)], 10) / (10 * K)
side = torch.cuda.Stream()
graph = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    for i in range(3):
        launch(i, side.cuda_stream)
    side.synchronize()
    with torch.cuda.graph(graph, stream=side):
        for i in range(K):
            launch(i, side.cuda_stream)
graph.replay()
us_graph = timed(graph.replay, 10) / (10 * K)
print(f"B={B} C={C} S={S}: stream {us_stream:.2f} us/launch, graph {us_graph:.2f} us/launch")
