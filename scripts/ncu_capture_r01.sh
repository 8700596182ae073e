cd /root/repo
cap() { # name skip
  timeout 300 ncu --set full --import-source on --clock-control none -s $2 -c 1 -f -o gpurun_out/$1 python scripts/ncu_target.py > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_r01_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/prof_$1_r01_details.txt 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
cap tc_fused_bf16_c192 222
cap tc_fused_bf16_c96 226
cap tcp_c64k7 116
cap tcp_c128k7 123
cap tcp_c512k7 137
cap lstm_dec_bf16 204
cap tc_tf32_c64k1 117
cap tc_bf16_c384k7 215
ls -la gpurun_out/prof_*_r01_raw.csv | wc -l
