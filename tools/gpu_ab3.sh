#!/bin/bash
# A/B of two library builds over the bench configurations: tools/gpu_ab3.sh <libA> <libB> [settings]
A=$1; B=$2; shift 2
for cb in "joint5 8192" "joint5 1024" "joint5 256" "swd2 4096" "transd3 4096" "transd3 1024" "transd3 256"; do
  set -- $cb
  for lib in $A $B; do
    BH_B200_LIB=bayhunter_b200/variants/libbh_$lib.so timeout 120 python tools/quick_bench.py $1 $2 swd_pool=-1 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    k = d.get('kernels', {})
    print('$1 $2 $lib total', d.get('total_ms'), 'swd', k.get('swd'), k.get('swd_pool'), 'consumed', d.get('consumed'), 'evaluated', d.get('evaluated'), 'logL', d.get('logL_sum'))
"
  done
done
