#!/bin/bash
# Round-end measurement pass on ONE B200 (run under gpurun from the repo root): GPU tests, the bench line,
# the per-configuration sweep, the ncu launch list of the bench command and one full capture of the fused
# loss kernel.  Everything lands in gpurun_out/ ; copy what is to be judged into profiles/.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2>> gpurun_out/bench_n1.err
python tools/configs_bench.py > gpurun_out/configs.log 2>&1; cat gpurun_out/configs.log
python tools/launch_gap.py > gpurun_out/launch_gap.log 2>&1; python tools/launch_gap.py 64 2 1 256 256 >> gpurun_out/launch_gap.log 2>&1
cat gpurun_out/launch_gap.log
# launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
python bench.py --steps 4 --warmup 3 --skip-cpu --e2e-steps 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --skip-cpu --e2e-steps 2 > gpurun_out/ncu_list.log 2>&1
# the dominant kernel, one full capture
python tools/prof_case.py f32 i64 8 2 128 128 128 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:loss_kernel -s 3 -c 1 -o gpurun_out/loss_f32 \
    python tools/prof_case.py f32 i64 8 2 128 128 128 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
