#!/bin/bash
# One 8-GPU box (gpurun --gpus 8): the 8-rank parity worker (incl. full-size cfg5 / cfg2 steps and the fused vote + gather),
# the loss bench line and the vote workload at N = 8.  Everything lands in gpurun_out/.
set -u
R=${1:-r02}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29701 tests/dist_worker.py cfg5 cfg2 > gpurun_out/${R}_dist_worker_n8.log 2>&1; grep -E "OK|FAIL|Error" gpurun_out/${R}_dist_worker_n8.log | tail -5
run 29702 bench.py --gpus 8 --skip-cpu > gpurun_out/${R}_bench_n8.json 2> gpurun_out/${R}_bench_n8.err; head -c 400 gpurun_out/${R}_bench_n8.json; echo
run 29703 bench.py --gpus 8 --workload vote --skip-cpu > gpurun_out/${R}_bench_vote_n8.json 2>> gpurun_out/${R}_bench_n8.err; head -c 300 gpurun_out/${R}_bench_vote_n8.json; echo
tail -c 300 gpurun_out/${R}_bench_n8.err
