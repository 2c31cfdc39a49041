#!/usr/bin/env python
"""Half-precision loss cases under the shell choices (DS_LOSS_TMA=-1/0/1), for A/B runs of library variants."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = sys.argv[:1]
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("configs_bench_mod", os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs_bench.py"))
src = open(spec.origin).read().split("\nif __name__")[0]
mod = {"__name__": "configs_bench_mod", "__file__": spec.origin}
exec(compile(src, spec.origin, "exec"), mod)
loss_case = mod["loss_case"]
f16, i64, u8 = torch.float16, torch.int64, torch.uint8
for tma in os.environ.get("SHELLS", "-1,1").split(","):
    os.environ["DS_LOSS_TMA"] = tma
    print(f"== DS_LOSS_TMA={tma}", flush=True)
    loss_case("cfg5 8x3x192x160x160 fp16/int64", 8, 3, (192, 160, 160), f16, i64, nsets=2, steps=50)
    loss_case("cfg5 8x3x192x160x160 fp16/int64 fwd only", 8, 3, (192, 160, 160), f16, i64, grad=False, nsets=3, steps=50)
    loss_case("cfg2 8x2x128^3 fp16/int64", 8, 2, (128, 128, 128), f16, i64, nsets=4, steps=200)
    loss_case("cfg2 8x2x128^3 fp16/int64 fwd only", 8, 2, (128, 128, 128), f16, i64, grad=False, nsets=4, steps=200)
    loss_case("cfg2 8x2x128^3 fp16/uint8", 8, 2, (128, 128, 128), f16, u8, nsets=6, steps=200)
    loss_case("cfg5 8x3x192x160x160 fp32/int64", 8, 3, (192, 160, 160), torch.float32, i64, nsets=2, steps=50)
