#!/bin/bash
# first GPU call of this session: parity tests + A/B of the dispersion kernel variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
V0="swd_searches_per_warp=32,swd_max_spec=8,concurrent=0 swd_searches_per_warp=16,swd_max_spec=8,concurrent=0 swd_searches_per_warp=16,swd_max_spec=8,concurrent=1"
BH_B200_LIB=$PWD/bayhunter_b200/variants/libbh_v0.so python tools/quick_bench.py joint5 8192 $V0 > gpurun_out/c1_ab_v0.log 2>&1
python tools/quick_bench.py joint5 8192 > gpurun_out/c1_ab_v1.log 2>&1
python tools/quick_bench.py swd2 4096 swd_searches_per_warp=8,swd_group_searches_per_warp=4,swd_max_spec=8 swd_searches_per_warp=4,swd_group_searches_per_warp=4,swd_max_spec=8 swd_searches_per_warp=4,swd_group_searches_per_warp=2,swd_max_spec=8 swd_searches_per_warp=2,swd_group_searches_per_warp=2,swd_max_spec=16 >> gpurun_out/c1_ab_v1.log 2>&1
python tools/quick_bench.py transd3 4096 swd_searches_per_warp=8,swd_group_searches_per_warp=8,swd_max_spec=8 swd_searches_per_warp=8,swd_group_searches_per_warp=4,swd_max_spec=8 swd_searches_per_warp=4,swd_group_searches_per_warp=2,swd_max_spec=8 >> gpurun_out/c1_ab_v1.log 2>&1
cat gpurun_out/c1_ab_v0.log gpurun_out/c1_ab_v1.log
