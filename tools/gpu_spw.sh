#!/bin/bash
# per-curve-type models-per-warp sweep (joint5, B=8192)
mkdir -p gpurun_out
python tools/quick_bench.py joint5 8192 \
  concurrent=1 \
  swd_spw_rg=8,concurrent=1 \
  swd_spw_rg=8,swd_spw_lp=32,concurrent=1 \
  swd_spw_rg=8,swd_spw_lp=32,swd_spw_rp=32,concurrent=1 \
  swd_spw_lp=32,concurrent=1 \
  swd_spw_lp=32,swd_spw_rp=32,concurrent=1 \
  swd_spw_rg=8,swd_spw_lg=8,swd_spw_lp=32,swd_spw_rp=32,concurrent=1 \
  swd_spw_rg=8,swd_spw_rp=8,concurrent=1 \
  swd_spw_rg=4,swd_spw_rp=8,swd_spw_lp=32,concurrent=1 \
  2>&1 | tee gpurun_out/spw_sweep.log | python tools/fmt_ab.py
