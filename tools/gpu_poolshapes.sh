#!/bin/bash
# pool kernel shapes: variants built by tools/build_variant.sh pw<warps>b<minblocks>
run() { lib=$1; shift; BH_B200_LIB=bayhunter_b200/variants/libbh_$lib.so timeout 120 python tools/quick_bench.py ${CFG:-joint5} ${NB:-8192} "$@" 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    k = d.get('kernels', {})
    print('$lib', {a: d[a] for a in d if a.startswith('swd_') or a.startswith('rf_')}, 'total', d.get('total_ms'), 'swd', k.get('swd'), 'love', k.get('swd_love'), 'evaluated', d.get('evaluated'), 'rounds', d.get('rounds'), 'same', d.get('same_as_first'))
"; }
"$@"
