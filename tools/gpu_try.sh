#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in default sm128 sm64 sm128mb5 rfmb5 rfmb6; do
  if [ "$v" = default ]; then unset BH_B200_LIB; else export BH_B200_LIB=$PWD/bayhunter_b200/variants/libbh_$v.so; fi
  python tools/quick_bench.py joint5 8192 concurrent=1 concurrent=0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); print('$v', 'conc', d['concurrent'], 'total', d['total_ms'], {k:round(v,3) for k,v in d['kernels'].items() if k.startswith('rf') or k=='swd'})
"
done
unset BH_B200_LIB
python tools/quick_bench.py transd3 4096 concurrent=1 concurrent=0 2>&1 | python tools/fmt_ab.py | grep -v lib | cut -c1-250
