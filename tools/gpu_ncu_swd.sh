#!/bin/bash
# ncu --set full capture of the dispersion kernels (first launch of each after one warm-up eval)
# usage: tools/gpu_ncu_swd.sh <out-name> <cfg> <B> <settings> [count]
mkdir -p gpurun_out
OUT=$1; CFG=$2; B=$3; SET=$4; CNT=${5:-2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:swd_kernel -s $CNT -c $CNT \
    -f -o gpurun_out/$OUT python tools/quick_bench.py $CFG $B $SET > gpurun_out/$OUT.log 2>&1
tail -3 gpurun_out/$OUT.log
