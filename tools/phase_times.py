#!/usr/bin/env python
"""Phase breakdown of the LDG loss kernel from a -DDS_TIMING build (tuning only).
    python -m deep_staple_b200.build --variant timing -DDS_TIMING
    DS_LIB_PATH=build/alt/timing/libdeep_staple_b200.so python tools/phase_times.py B C D H W
Stamps per CTA (thread 0, %globaltimer ns): 0 entry, 1 prologue done, 2 first step done, 3 loop done,
4 statistics flushed, 5 fence + barrier, 6 ticket known, 7 epilogue done (last CTA only)."""
import ctypes
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_staple_b200 import _binding as b  # noqa: E402

B, C = int(sys.argv[1]), int(sys.argv[2])
S = tuple(int(x) for x in sys.argv[3:6])
V = math.prod(S)
dev = torch.device("cuda:0")
lib = b.load()
raw = ctypes.CDLL(b.lib_path())
sets = []
for i in range(3):
    g = torch.Generator(device=dev).manual_seed(i)
    sets.append((2 * torch.randn((B, C, V), generator=g, device=dev),
                 (torch.rand((B, V), generator=g, device=dev) < 0.05).long(),
                 torch.empty((B, C, V), device=dev)))
bare = torch.randn(B, device=dev); fixed = torch.rand(B, device=dev) + 1; cw = torch.ones(C, device=dev)
scale = torch.tensor(1024.0, device=dev); stats = torch.zeros((B, C, 4), dtype=torch.float64, device=dev)
loss_out = torch.zeros(2, device=dev); d_bare = torch.zeros(B, device=dev); per_inst = torch.zeros((B, 4), device=dev)
ws = torch.zeros((int(lib.ds_dp_loss_workspace_bytes(B, C)) + 7) // 8, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream


def launch(i):
    z, t, g = sets[i % 3]
    rc = lib.ds_dp_loss_fwd_bwd(z.data_ptr(), b.DS_F32, t.data_ptr(), b.DS_I64, None, B, C, V, bare.data_ptr(),
                                fixed.data_ptr(), B, 0, cw.data_ptr(), 1, scale.data_ptr(), g.data_ptr(),
                                stats.data_ptr(), loss_out.data_ptr(), d_bare.data_ptr(), per_inst.data_ptr(),
                                ws.data_ptr(), ws.numel() * 8, stream)
    assert rc == 0


for i in range(20):
    launch(i)
torch.cuda.synchronize()
buf = np.zeros((2048, 16), dtype=np.uint64)
rows = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(4):
        launch(i)          # keep the pipeline busy: the measured launch is the last of five back to back
    e0.record()
    launch(4)
    e1.record()
    torch.cuda.synchronize()
    raw.ds_debug_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes))
    t = buf.astype(np.int64)
    n = int((t[:, 0] > 0).sum())
    t = t[:n]
    t0 = t[:, 0].min()
    last = int(np.argmax(t[:, 7]))
    rel = (t - t0) / 1000.0
    names = ["entry", "prologue", "step1", "loop", "flush", "fence", "ticket"]
    line = f"grid {n} event {e0.elapsed_time(e1) * 1000:.1f}us | "
    for k, nm in enumerate(names):
        line += f"{nm} {np.min(rel[:, k]):.1f}/{np.median(rel[:, k]):.1f}/{np.max(rel[:, k]):.1f} "
    line += f"| last CTA {last}: ticket {rel[last, 6]:.1f} end {rel[last, 7]:.1f} | fill {np.median(rel[:, 9]):.2f} pro {np.median(rel[:, 1]):.2f} | last: loop {rel[last, 3]:.2f} flush {rel[last, 4]:.2f} fence/records {rel[last, 5]:.2f} pre {rel[last, 10]:.2f} ticket {rel[last, 6]:.2f} fence2/polled {rel[last, 11]:.2f} stats {rel[last, 12]:.2f} end {rel[last, 7]:.2f} | slowest CTA records/fence {np.max(rel[:, 5]):.2f}"
    print(line)
    if rep == 4 and os.environ.get("DS_DUMP"):
        np.save(os.environ["DS_DUMP"], t)
