#!/bin/bash
export BH_B200_LIB=$PWD/bayhunter_b200/variants/libbh_exptab.so
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
unset BH_B200_LIB
bash tools/gpu_ab2.sh "default exptab" joint5 8192 2>&1 | python tools/fmt_ab.py | cut -c1-260
