#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py joint5 8192 concurrent=1 concurrent=0 concurrent=1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); print('conc', d['concurrent'], 'total', d['total_ms'], {k:round(v,3) for k,v in d['kernels'].items()}, 'logL_sum', d['logL_sum'])
"
python tools/quick_bench.py transd3 4096 concurrent=1 2>&1 | python tools/fmt_ab.py | grep -v lib | cut -c1-250
