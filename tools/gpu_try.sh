#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/autotune_probe.py joint5 2048 4096 6144 8192 12288
python tools/autotune_probe.py transd3 4096
python tools/autotune_probe.py swd2 4096
