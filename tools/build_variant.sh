#!/bin/bash
# tools/build_variant.sh <name> [extra nvcc flags...]  -> bayhunter_b200/variants/libbh_<name>.so (working tree sources)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p bayhunter_b200/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -cudart static "$@" \
  -o bayhunter_b200/variants/libbh_$NAME.so bayhunter_b200/csrc/{engine,prep_kernel,swd_kernel,swd_lockstep,swd_pool,swd_general,rf_kernel,loglik_kernel,sampler,noise_kernel}.cu
echo built bayhunter_b200/variants/libbh_$NAME.so
