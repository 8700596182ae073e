#!/bin/bash
# Runs scripts/_bin/mma_probe over the config grid (one process per config, bounded by timeout); log -> gpurun_out/mma_probe.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=scripts/_bin/mma_probe
LOG=gpurun_out/mma_probe.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $LOG
run() { timeout 30 $B "$@" >> $LOG 2>&1 || echo "config $* failed rc=$?" >> $LOG; }
#    kind N nacc run issuers occ ts cg [b_period]
# raw rate vs N, one accumulator chain
for N in 32 64 128 256; do run 0 $N 1 336 1 1 0 1 6; done
for N in 32 64 128 256; do run 1 $N 1 336 1 1 0 1 6; done
# accumulator rotation: run length 1 / 3 / 6
run 0 64 4 1 1 1 0 1 12
run 0 64 4 3 1 1 0 1 12
run 0 64 4 6 1 1 0 1 24
run 0 128 2 1 1 1 0 1 6
run 0 128 2 3 1 1 0 1 6
run 0 128 4 1 1 1 0 1 12
run 0 256 2 1 1 1 0 1 6
run 0 256 2 3 1 1 0 1 6
run 1 64 4 6 1 1 0 1 24
run 1 64 4 1 1 1 0 1 24
run 1 128 2 6 1 1 0 1 12
run 1 128 2 1 1 1 0 1 12
run 1 256 2 1 1 1 0 1 12
# two issuing warps in one CTA
run 0 64 2 3 2 1 0 1 6
run 0 128 2 3 2 1 0 1 6
run 0 128 1 336 2 1 0 1 6
run 1 64 2 6 2 1 0 1 12
run 1 128 2 6 2 1 0 1 12
# two CTAs per SM
run 0 64 2 3 1 2 0 1 6
run 0 128 2 3 1 2 0 1 6
run 0 128 1 336 1 2 0 1 6
run 1 64 2 6 1 2 0 1 12
run 1 128 2 6 1 2 0 1 12
# A from TMEM
for N in 32 64 128 256; do run 0 $N 1 336 1 1 1 1 6; done
run 0 128 2 1 1 1 1 1 6
run 0 256 1 336 1 1 1 1 1
run 1 128 2 1 1 1 1 1 6
run 1 256 1 336 1 1 1 1 6
run 0 128 1 336 2 1 1 1 6
# cta_group::2 (M = 256 over an SM pair)
run 0 64 2 3 1 1 0 2 6
run 0 128 2 3 1 1 0 2 6
run 0 256 2 3 1 1 0 2 6
run 0 256 1 336 1 1 0 2 6
run 1 128 2 6 1 1 0 2 12
run 1 256 2 6 1 1 0 2 12
cat $LOG
