#!/bin/bash
# ncu --set full capture of the dispersion kernel (one launch) for a given setting
# usage: tools/gpu_ncu_swd.sh <out-name> <cfg> <B> <settings>
mkdir -p gpurun_out
OUT=$1; CFG=$2; B=$3; SET=$4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swd_kernel -s 1 -c 1 \
    -f -o gpurun_out/$OUT python tools/quick_bench.py $CFG $B $SET > gpurun_out/$OUT.log 2>&1
tail -3 gpurun_out/$OUT.log
