#!/bin/bash
# the pool kernel below its rule's threshold, with the deep guess trees (current build)
run() { cfg=$1; B=$2; shift 2; S="swd_pool=0"; for m in "$@"; do S="$S swd_pool=1,swd_pool_models=$m"; done
  timeout 200 python tools/quick_bench.py $cfg $B $S 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    k = d['kernels']
    print('$cfg $B', 'pool M=%s' % d.get('swd_pool_models') if d.get('swd_pool') else 'swd_kernel', 'total', d.get('total_ms'), 'swd', k.get('swd'), k.get('swd_pool'), k.get('swd_pool_love'), 'evaluated', d.get('evaluated'), 'same', d.get('same_as_first'))
"; }
run joint5 3072 11 12
run joint5 2048 7 8
run joint5 1024 4
run swd2 4096 7 14
run swd2 8192 14
run transd3 2048 4 8
