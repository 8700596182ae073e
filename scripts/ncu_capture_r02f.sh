#!/bin/bash
# Round-2 FINAL ncu evidence (run under gpurun, one GPU): launch list of two full B=32 x 4 s forwards + full captures of the
# kernels changed after the first round-2 capture set (conv_tc_kernel's slot-structured issue loop / 16-worker tiles) and of
# the unchanged heavy hitters for the same build.  Exports land in gpurun_out/; the .ncu-rep files are deleted (size).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r02f.csv python scripts/ncu_target.py > gpurun_out/ncu_list_r02f.log 2>&1
cap() { # name kernel-regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_r02f_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/prof_${name}_r02f_details.txt 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap tc_fused_c192 conv_tc_kernel 1 python scripts/gpu_resunit_one.py 192 3 48000 6
cap tc_fused_c96 conv_tc_kernel 1 python scripts/gpu_resunit_one.py 96 3 96000 6
cap tc_c384k7 conv_tc_kernel 1 python scripts/gpu_tt_one.py 384 7 3 9600 5
cap tt_c128k7 conv_tt 1 python scripts/gpu_tt_one.py 128 7 3 48000
cap lstm2_enc lstm_rec2 2 python scripts/gpu_lstm_one.py 1024 0
ls -la gpurun_out/prof_*_r02f_raw.csv | wc -l
