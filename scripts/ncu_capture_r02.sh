#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU): launch list of two full B=32 x 4 s forwards + full captures of the kernels
# new in this round.  Exports (raw CSV + details text) land in gpurun_out/; the .ncu-rep files are deleted (size).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r02.csv python scripts/ncu_target.py > gpurun_out/ncu_list_r02.log 2>&1
cap() { # name kernel-regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_r02_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/prof_${name}_r02_details.txt 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap tt_c128k7 conv_tt 1 python scripts/gpu_tt_one.py 128 7 3 48000
cap tt_c64k7 conv_tt 1 python scripts/gpu_tt_one.py 64 7 3 96000
cap tt_c512k7 conv_tt 1 python scripts/gpu_tt_one.py 512 7 1 1920
cap tt_c256k1 conv_tt 1 python scripts/gpu_tt_one.py 256 1 1 9600
cap lstm2_enc lstm_rec2 2 python scripts/gpu_lstm_one.py 1024 0
cap lstm2_dec lstm_rec2 2 python scripts/gpu_lstm_one.py 1536 1
cap rvq rvq_kernel 1 python scripts/gpu_rvq_one.py
ls -la gpurun_out/prof_*_r02_raw.csv | wc -l
