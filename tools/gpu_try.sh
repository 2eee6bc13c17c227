#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py joint5 8192 concurrent=1 concurrent=0 concurrent=1 2>&1 | python tools/fmt_ab.py | cut -c1-330
python tools/quick_bench.py swd2 4096 concurrent=1 2>&1 | python tools/fmt_ab.py | cut -c1-330
