#!/bin/bash
# models per warp (phase / group) at small batches, with the deeper guess trees in the kernel
run() { cfg=$1; B=$2; shift 2; S="swd_pool=-1"; for pg in "$@"; do S="$S swd_searches_per_warp=${pg%/*},swd_group_searches_per_warp=${pg#*/}"; done
  timeout 200 python tools/quick_bench.py $cfg $B $S 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$cfg $B', d.get('swd_searches_per_warp', 'rule'), d.get('swd_group_searches_per_warp', ''), 'total', d.get('total_ms'), 'swd', d['kernels'].get('swd'), 'evaluated', d.get('evaluated'), 'same', d.get('same_as_first'))
"; }
run joint5 128 2/1 1/1 2/2
run joint5 256 4/2 2/2 2/1 1/1
run joint5 512 4/4 4/2 2/2 2/1
run joint5 1024 4/4 4/2 2/2
run swd2 1024 4/2 2/2 2/1 1/1
run transd3 1024 4/2 2/2 2/1
