#!/usr/bin/env python
"""Run a few launches of one loss configuration through the C ABI (for ncu).
usage: prof_case.py <f32|f16|bf16> <i64|u8> B C D H W [n] [grad=1]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_staple_b200 import _binding as b  # noqa: E402

ldt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[sys.argv[1]]
tdt = {"i64": torch.int64, "u8": torch.uint8}[sys.argv[2]]
B, C = int(sys.argv[3]), int(sys.argv[4])
S = tuple(int(x) for x in sys.argv[5:8])
n = int(sys.argv[8]) if len(sys.argv) > 8 else 6
grad = (int(sys.argv[9]) if len(sys.argv) > 9 else 1) != 0
DT = {torch.float32: b.DS_F32, torch.float16: b.DS_F16, torch.bfloat16: b.DS_BF16, torch.int64: b.DS_I64, torch.uint8: b.DS_U8}
dev = torch.device("cuda:0")
lib = b.load()
V = math.prod(S)
sets = []
for i in range(3):
    g = torch.Generator(device=dev).manual_seed(i)
    z = (2 * torch.randn((B, C, V), generator=g, device=dev)).to(ldt)
    t = (torch.rand((B, V), generator=g, device=dev) < 0.05).to(tdt)
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
for i in range(n):
    z, t, g = sets[i % 3]
    rc = lib.ds_dp_loss_fwd_bwd(z.data_ptr(), DT[ldt], t.data_ptr(), DT[tdt], None, B, C, V, bare.data_ptr(),
                                fixed.data_ptr(), B, 0, cw.data_ptr(), 1, scale.data_ptr(),
                                g.data_ptr() if grad else None, stats.data_ptr(), loss_out.data_ptr(),
                                d_bare.data_ptr(), per_inst.data_ptr(), ws.data_ptr(), ws.numel() * 8, stream)
    assert rc == 0, lib.ds_last_error_string()
torch.cuda.synchronize()
print("done", sys.argv[1:], float(loss_out[0]))
