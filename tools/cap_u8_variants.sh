cap() {
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/r02_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  if [ -f gpurun_out/r02_$name.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/r02_$name.ncu-rep > gpurun_out/r02_${name}_ncu_summary.txt 2>&1
    ncu -i gpurun_out/r02_$name.ncu-rep --page source --csv > /tmp/ncu_src_$name.csv 2>/dev/null
    python tools/ncu_source_hot.py /tmp/ncu_src_$name.csv 30 > gpurun_out/r02_${name}_ncu_source_hot.txt 2>&1
    rm -f gpurun_out/r02_$name.ncu-rep
  fi
}
cap loss_f32_u8 loss_kernel 3 python tools/prof_case.py f32 u8 8 2 128 128 128
cap loss_f16_u8 loss_kernel 3 python tools/prof_case.py f16 u8 8 2 128 128 128
cap loss_f16_c2 loss_kernel 3 python tools/prof_case.py f16 i64 8 2 128 128 128
head -20 gpurun_out/r02_loss_f32_u8_ncu_summary.txt
