#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py joint5 8192 concurrent=1 swd_split_waves=1,concurrent=1 swd_split_waves=1,rf_after_love=1,concurrent=1 swd_split_waves=0,rf_after_love=0,concurrent=1 2>&1 | python tools/fmt_ab.py | cut -c1-330
