#!/usr/bin/env python
"""Launch-time floor of each kernel family on a tiny problem (what does NOT shrink with the data)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import deep_staple_b200 as ds
from deep_staple_b200 import _binding as b
dev = torch.device("cuda:0"); lib = b.load()
B, C, V = 8, 2, 4096
z = torch.randn((B, C, V), device=dev); t = (torch.rand((B, V), device=dev) < 0.3).long(); g = torch.empty_like(z)
bare = torch.randn(B, device=dev); cw = torch.ones(C, device=dev); scale = torch.tensor(1024.0, device=dev)
stats = torch.zeros((B, C, 4), dtype=torch.float64, device=dev); loss_out = torch.zeros(2, device=dev)
d_bare = torch.zeros(B, device=dev); per_inst = torch.zeros((B, 4), device=dev)
ws = torch.zeros((int(lib.ds_dp_loss_workspace_bytes(B, C)) + 7) // 8, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
w = torch.ones(B, device=dev)
empty = torch.empty(1, device=dev)
cases = {
  "torch empty-ish kernel (fill_ 1 elem)": lambda: empty.fill_(1.0),
  "bwd-only (no statistics, no epilogue)": lambda: lib.ds_dp_loss_backward_logits(z.data_ptr(), b.DS_F32, t.data_ptr(), b.DS_I64, B, C, V, w.data_ptr(), 1, scale.data_ptr(), g.data_ptr(), st),
  "fwd+bwd, finalize": lambda: lib.ds_dp_loss_fwd_bwd(z.data_ptr(), b.DS_F32, t.data_ptr(), b.DS_I64, None, B, C, V, bare.data_ptr(), None, B, 0, cw.data_ptr(), 1, scale.data_ptr(), g.data_ptr(), stats.data_ptr(), loss_out.data_ptr(), d_bare.data_ptr(), per_inst.data_ptr(), ws.data_ptr(), ws.numel()*8, st),
  "fwd+bwd, rows only (B_total = 2B: no finalize)": lambda: lib.ds_dp_loss_fwd_bwd(z.data_ptr(), b.DS_F32, t.data_ptr(), b.DS_I64, None, B, C, V, torch.cat([bare,bare]).data_ptr() if False else bare2.data_ptr(), None, 2*B, 0, cw.data_ptr(), 1, scale.data_ptr(), g.data_ptr(), stats2.data_ptr(), loss_out.data_ptr(), d_bare2.data_ptr(), per_inst2.data_ptr(), ws.data_ptr(), ws.numel()*8, st),
  "finalize kernel alone": lambda: lib.ds_dp_loss_finalize(stats.data_ptr(), B, C, V, bare.data_ptr(), None, cw.data_ptr(), 1, scale.data_ptr(), loss_out.data_ptr(), d_bare.data_ptr(), per_inst.data_ptr(), st),
  "label_stats": lambda: lib.ds_label_stats(t.data_ptr(), b.DS_I64, t.data_ptr(), b.DS_I64, B, C, V, stats.data_ptr(), ws.data_ptr(), ws.numel()*8, st),
}
bare2 = torch.randn(2*B, device=dev); stats2 = torch.zeros((2*B, C, 4), dtype=torch.float64, device=dev)
d_bare2 = torch.zeros(2*B, device=dev); per_inst2 = torch.zeros((2*B, 4), device=dev)
for name, fn in cases.items():
    sec = bench.timed(fn, 2000, 50) / 2000
    print(f"{name:50s} {sec*1e6:7.2f} us")
