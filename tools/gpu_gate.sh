#!/bin/bash
# RF gate threshold sweep under bench conditions (BH_GATE overrides the default 25 %)
for g in 5 15 25 40 60 80; do
BH_GATE=$g python bench.py --steps 40 --warmup 3 --no-cpu-baseline --sampler-iters 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('gate', '$g', 'ms_per_step', round(d['ms_per_step'],3), 'e2e ms', round(8192e3/d['e2e']['value'],3), {k:round(v,2) for k,v in d['kernel_ms'].items()})
"
done
