#!/bin/bash
mkdir -p gpurun_out
python tools/quick_bench.py joint5 8192 concurrent=1 rf_first=1,concurrent=1 rf_first=0,concurrent=1 rf_first=1,concurrent=1 2>&1 | python tools/fmt_ab.py | cut -c1-330
python tools/quick_bench.py transd3 4096 concurrent=1 rf_first=1,concurrent=1 2>&1 | python tools/fmt_ab.py | cut -c1-330
