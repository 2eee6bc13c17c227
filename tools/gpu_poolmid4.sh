#!/bin/bash
source /dev/null
run() { cfg=$1; B=$2; shift 2; S="swd_pool=0"; for m in "$@"; do S="$S swd_pool=1,swd_pool_models=$m"; done
  timeout 200 python tools/quick_bench.py $cfg $B $S 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    k = d['kernels']
    print('$cfg $B', 'pool M=%s' % d.get('swd_pool_models') if d.get('swd_pool') else 'swd_kernel', 'total', d.get('total_ms'), 'swd', k.get('swd'), k.get('swd_pool'), k.get('swd_pool_love'), 'evaluated', d.get('evaluated'), 'same', d.get('same_as_first'))
"; }
run swd2 16384 21 28 32
run swd2 2048 7 14
run transd3 1024 4 8
run joint5 2560 9 10 14
run swd2 6144 14 16
