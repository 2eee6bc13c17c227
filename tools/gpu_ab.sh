#!/bin/bash
# tools/gpu_ab.sh <cfg> <B> "<variant names>" <settings...>: quick_bench for each library variant ("default" = in-tree build)
mkdir -p gpurun_out
CFG=$1; B=$2; VARS=$3; shift 3
for v in $VARS; do
  if [ "$v" = default ]; then unset BH_B200_LIB; else export BH_B200_LIB=$PWD/bayhunter_b200/variants/libbh_$v.so; fi
  python tools/quick_bench.py $CFG $B "$@" 2>&1 | tee -a gpurun_out/ab_$CFG.log
done
