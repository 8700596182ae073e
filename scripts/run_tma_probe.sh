#!/bin/bash
# Runs scripts/_bin/tma_probe (nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/tma_probe.cu): unicast vs cluster-multicast
# streaming of one L2-resident weight region by every SM; log -> gpurun_out/tma_probe.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=scripts/_bin/tma_probe
LOG=gpurun_out/tma_probe.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $LOG
run() { timeout 30 $B "$@" >> $LOG 2>&1 || echo "config $* failed rc=$?" >> $LOG; }
# 516 KB region = the fused C = 192 unit's fp16 k = 7 weights; 129 KB = C = 96; slots of 42 KB / 21 KB / 8 KB
for C in 1 2 4; do
  run $C 43008 516096 600
  run $C 21504 129024 1200
  run $C 8192 524288 3000
  run $C 43008 516096 600 74      # half the SMs: per-SM ingest limit vs chip limit
done
cat $LOG
