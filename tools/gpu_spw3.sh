#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/autotune_probe.py joint5 2048 4096 6144 7168 7936 8192 9216 12288 16384
python tools/quick_bench.py joint5 8192 concurrent=1 concurrent=0 concurrent=1 2>&1 | python tools/fmt_ab.py | grep -v lib | cut -c1-260
