#!/bin/bash
# ragged batch (transd3: 3..31 rows): models per warp (phase, group) vs the rule (0 0)
BS=${1:-4096}
ARGS=""
for q in "0 0" "16 8" "8 8" "8 4" "4 4" "4 2"; do set -- $q; ARGS="$ARGS swd_searches_per_warp=$1,swd_group_searches_per_warp=$2,concurrent=1"; done
for B in $BS; do
python tools/quick_bench.py transd3 $B $ARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l)
    if 'total_ms' not in d: print(l.strip()[:200]); continue
    print('B',d['B'],'S',d['swd_searches_per_warp'],'Sg',d['swd_group_searches_per_warp'],'total',d['total_ms'],'swd',round(d['kernels']['swd'],2),'rounds(max)',d['rounds'][1:4:2],d['same_as_first'])
"
done
