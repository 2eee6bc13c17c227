#!/bin/bash
# tools/gpu_ab2.sh "<variants>" [cfg] [B]: quick_bench with the engine's default tunables, 2 passes per variant
mkdir -p gpurun_out
VARS=$1; CFG=${2:-joint5}; B=${3:-8192}
for pass in 1 2; do
for v in $VARS; do
  if [ "$v" = default ]; then unset BH_B200_LIB; else export BH_B200_LIB=$PWD/bayhunter_b200/variants/libbh_$v.so; fi
  python tools/quick_bench.py $CFG $B concurrent=1 concurrent=0 2>&1 | tee -a gpurun_out/ab2_$CFG.log
done; done
