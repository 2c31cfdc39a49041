#!/usr/bin/env python
"""Run a few launches of one loss-kernel variant (for ncu).  usage: prof_step.py fwdbwd|bwd|fwd [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deep_staple_b200 import _binding as b  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwdbwd"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
step = bench.LossStep(dev, 3, 0, 1)
lib = b.load()
B, C, V = step.B, step.C, step.V
w = torch.rand(B, device=dev) + 0.5
for i in range(n):
    s = step.sets[i % 3]
    if which == "fwdbwd":
        step.kernel()
    elif which == "bwd":
        lib.ds_dp_loss_backward_logits(s["logits"].data_ptr(), b.DS_F32, s["target"].data_ptr(), b.DS_I64, B, C, V,
                                       w.data_ptr(), 1, step.scale.data_ptr(), step.dlogits[i % 3].data_ptr(), step.stream)
    else:
        lib.ds_dp_loss_fwd_bwd(s["logits"].data_ptr(), b.DS_F32, s["target"].data_ptr(), b.DS_I64, None, B, C, V,
                               step.bare.data_ptr(), step.fixed.data_ptr(), B, 0, step.cw.data_ptr(), 1,
                               step.scale.data_ptr(), None, step.stats.data_ptr(), step.loss_out.data_ptr(),
                               step.d_bare.data_ptr(), step.per_inst.data_ptr(), step.ws.data_ptr(), step.ws.numel() * 8,
                               step.stream)
torch.cuda.synchronize()
print("done", which, n)
