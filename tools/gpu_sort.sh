#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py transd3 4096 swd_sort_layers=0,concurrent=1 swd_sort_layers=1,concurrent=1 swd_sort_layers=1,swd_searches_per_warp=16,swd_group_searches_per_warp=8,concurrent=1 2>&1 | python tools/fmt_ab.py
python tools/quick_bench.py joint5 8192 swd_sort_layers=0,concurrent=1 swd_sort_layers=1,concurrent=1 2>&1 | python tools/fmt_ab.py
