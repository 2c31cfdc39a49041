#!/bin/bash
# Round-end measurement pass on ONE B200 (run under gpurun from the repo root): GPU tests, the bench line, the
# reference arm, the per-configuration sweep, host overhead of the Python layer, the ncu launch list of the bench command
# and full captures of the dominant kernels.  Everything lands in gpurun_out/ ; copy what is to be judged into profiles/.
set -u
R=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${R}_pytest_gpu.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; head -c 600 gpurun_out/${R}_bench_n1.json; echo
python bench.py --workload vote > gpurun_out/${R}_bench_vote_n1.json 2>> gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${R}_bench_reference_arm.json 2>> gpurun_out/${R}_bench_n1.err
python tools/configs_bench.py > gpurun_out/${R}_configs.log 2>&1; cat gpurun_out/${R}_configs.log
python tools/host_overhead.py > gpurun_out/${R}_host_overhead.log 2>&1; cat gpurun_out/${R}_host_overhead.log
# launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
python bench.py --steps 4 --warmup 3 --skip-cpu --e2e-steps 2 --skip-configs > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 4 --warmup 3 --skip-cpu --e2e-steps 2 --skip-configs > gpurun_out/ncu_list.log 2>&1
python tools/launches_summary.py gpurun_out/${R}_launches.csv > gpurun_out/${R}_launches_table.txt 2>&1
# the dominant kernels, one full capture each; the reports are read ON the box (a report with imported sources is
# ~50 MB, gpurun_out/ carries 64 MB back): raw metrics + summary + per-line source view as text, then the report goes
cap() {  # cap <name> <kernel regex> <skip> <command...>
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/${R}_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  if [ -f gpurun_out/${R}_$name.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/${R}_$name.ncu-rep > gpurun_out/${R}_${name}_ncu_summary.txt 2>&1
    ncu -i gpurun_out/${R}_$name.ncu-rep --page raw --csv > gpurun_out/${R}_${name}_ncu_raw.csv 2>/dev/null
    ncu -i gpurun_out/${R}_$name.ncu-rep --page source --csv > /tmp/ncu_src_$name.csv 2>/dev/null
    python tools/ncu_source_hot.py /tmp/ncu_src_$name.csv 30 > gpurun_out/${R}_${name}_ncu_source_hot.txt 2>&1
    rm -f gpurun_out/${R}_$name.ncu-rep
  fi
}
cap loss_f32 loss_kernel 3 python tools/prof_case.py f32 i64 8 2 128 128 128
cap loss_f16_c3 loss_kernel 3 python tools/prof_case.py f16 i64 8 3 192 160 160
cap loss_f32_c3 loss_kernel 3 python tools/prof_case.py f32 i64 8 3 192 160 160
cap vote_u8 vote_kernel 2 python tools/prof_vote.py u8 13 30
cap vote_i64 vote_kernel 2 python tools/prof_vote.py i64 13 30
rm -f gpurun_out/plain*.log
# lower-priority logs last: a call cut short by its time limit keeps everything above
python tools/host_pack_bench.py > gpurun_out/${R}_host_pack_bench.log 2>&1
python tools/tune_vote_staged.py > gpurun_out/${R}_vote_u8_staged_tune.log 2>&1
python tools/launch_gap.py > gpurun_out/${R}_launch_gap.log 2>&1; python tools/launch_gap.py 64 2 1 256 256 >> gpurun_out/${R}_launch_gap.log 2>&1
ls -la gpurun_out | tail -40
