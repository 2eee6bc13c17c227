#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for g in 0 10 25 40 60; do
BH_GATE=$g python bench.py --steps 30 --warmup 3 --no-cpu-baseline --sampler-iters 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('gate', '$g', 'ms_per_step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), {k:round(v,2) for k,v in d['kernel_ms'].items()})
"
done
