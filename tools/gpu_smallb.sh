#!/bin/bash
# small batches: models per warp (phase, group) vs the rule (0 0); the machine is far from full, the chains are the critical path
for B in 256 512 1024 2048; do
ARGS=""
for q in "0 0" "16 8" "8 4" "4 2" "2 1" "1 1"; do set -- $q; ARGS="$ARGS swd_searches_per_warp=$1,swd_group_searches_per_warp=$2,concurrent=1"; done
python tools/quick_bench.py ${CFG:-joint5} $B $ARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l)
    if 'total_ms' not in d: print(l.strip()[:200]); continue
    print('B',d['B'],'S',d['swd_searches_per_warp'],'Sg',d['swd_group_searches_per_warp'],'total',d['total_ms'],'swd',round(d['kernels']['swd'],2),'rounds(max)',d['rounds'][1:8:2],'evaluated/consumed',round(d['evaluated']/max(1,d['consumed']),2),d['same_as_first'])
"
done
