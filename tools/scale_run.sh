#!/bin/bash
# 1/2/4/8-GPU bench lines on one 8-GPU box (gpurun --gpus 8): loss step (weak scaling, fused exchange) and vote shards.
set -u
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --skip-cpu --skip-e2e > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --skip-cpu --skip-e2e > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  cat gpurun_out/scale_n$n.json
done
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -2
