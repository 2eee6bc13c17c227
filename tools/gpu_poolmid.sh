#!/bin/bash
# the pool kernel with refinement guesses at mid batch sizes: swd_kernel (rule) vs pool with M models per CTA
export BH_B200_LIB=bayhunter_b200/variants/libbh_poolg.so
run() { cfg=$1; B=$2; shift 2; S="swd_pool=0"; for m in "$@"; do S="$S swd_pool=1,swd_pool_models=$m"; done
  timeout 200 python tools/quick_bench.py $cfg $B $S 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    k = d['kernels']
    print('$cfg $B', 'pool M=%s' % d.get('swd_pool_models') if d.get('swd_pool') else 'swd_kernel', 'total', d.get('total_ms'), 'swd', k.get('swd'), k.get('swd_pool'), k.get('swd_pool_love'), 'evaluated', d.get('evaluated'), 'same', d.get('same_as_first'))
"; }
run joint5 6144 21 24 28
run joint5 4096 14 16 12
run joint5 3072 11 12 9
run joint5 2048 7 8 6
run joint5 1024 4 3 5
run joint5 512 2 3
